"""TEST INFRASTRUCTURE: a few iterations of the reference's DP batch loop (adorym/ptychography.py:783-1271,
update_scheme='immediate', single rank) evaluated with the oracle, and the synthetic-data recipe of
adorym/simulation.py:659-671 (predict with the true object, store magnitudes)."""
import numpy as np
import torch
from oracle import multislice_oracle as orc


def oracle_reconstruct(truth, max_batches, obj_size, probe_pos, theta_st, theta_end, n_theta, energy_ev, psize_cm,
                       free_prop_cm, minibatch_size, n_epochs, alpha_d, alpha_b, gamma, learning_rate, probe_type,
                       probe_mag_sigma, probe_phase_sigma, probe_phase_max, initial_guess, dtype=torch.float32):
    Y, X, Z = obj_size
    P = [72, 72]
    theta_ls = np.linspace(theta_st, theta_end, n_theta, dtype='float32')
    pos = np.round(np.array(probe_pos)).astype(int)
    voxel_nm = np.array([psize_cm] * 3) * 1e7
    h = orc.get_kernel(voxel_nm[-1], 1240. / energy_ev, voxel_nm, P)
    cfg = dict(this_obj_size=[Y, X, Z], probe_size=P, energy_ev=energy_ev, psize_cm=psize_cm, h=h,
               free_prop_cm=free_prop_cm, unknown_type='delta_beta')
    pr, pi = orc.gaussian_probe(P, 1, probe_mag_sigma, probe_phase_max, probe_phase_sigma)
    pr = torch.tensor(pr[None], dtype=dtype)
    pi = torch.tensor(pi[None], dtype=dtype)
    luts = [orc.rotation_coords([Y, X, Z], th) for th in theta_ls]
    tr = torch.tensor(truth, dtype=dtype)
    data = np.zeros([n_theta, len(pos), P[0], P[1]], dtype='float32')
    with torch.no_grad():
        for i in range(n_theta):
            data[i] = orc.predict_ptycho(tr, pr, pi, pos, luts[i], cfg).numpy()
    obj = torch.tensor(np.stack(initial_guess, -1), dtype=dtype)
    m = torch.zeros_like(obj)
    v = torch.zeros_like(obj)
    reg = dict(alpha_d=alpha_d, alpha_b=alpha_b, gamma=gamma)
    hist = []
    chunks = orc.epoch_schedule(0, n_theta, len(pos), minibatch_size, 1)
    i_opt = 0
    for i_batch in range(min(max_batches, len(chunks))):
        i_theta, ind = orc.rank_slice_of_chunk(chunks, i_batch, 0, minibatch_size, 1)
        loss, _, _, grads = orc.ptycho_loss_and_grads(obj, pr, pi, pos[ind], torch.tensor(data[i_theta, ind], dtype=dtype),
                                                      luts[i_theta], cfg, reg=reg, want=('obj',))
        obj, m, v = orc.adam_step(obj, grads['obj'], m, v, i_opt, step_size=learning_rate)
        hist.append(float(loss))
        if i_batch == len(chunks) - 1 or chunks[i_batch + 1][0, 0] != i_theta:
            i_opt += 1
    return data, hist, obj.numpy()

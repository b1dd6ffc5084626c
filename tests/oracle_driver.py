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


def oracle_reconstruct_full(data_abs, mask, obj_size, probe_pos, theta_st, theta_end, n_theta, energy_ev, psize_cm,
                            free_prop_cm, minibatch_size, n_epochs, alpha_d, alpha_b, gamma, reweighted_l1, learning_rate,
                            update_scheme, probe_mag_sigma, probe_phase_sigma, probe_phase_max, optimize_probe,
                            probe_learning_rate, optimize_all_probe_pos, all_probe_pos_learning_rate, non_negativity,
                            initial_guess, dtype=torch.float32, rotate_out_of_loop=False, **unused):
    """The reference's DP batch loop (adorym/ptychography.py:783-1271, single rank) with the options of the driver
    goldens: reweighted L1 weights refreshed every 10th batch (:995-999), gradient accumulation per update_scheme
    (:1063-1099), Adam with the per-angle step index (:1125,1267-1271; reset every epoch :848), non-negativity (:1138),
    update_parameters (optimizers.py:993-1089: probe Adam on the stacked [M,P,P,2] array, position Adam + drift removal),
    finite-support mask (:1210-1215).  Returns (loss history, obj, probe_real, probe_imag, probe_pos_correction)."""
    Y, X, Z = obj_size
    P = [72, 72]
    theta_ls = np.linspace(theta_st, theta_end, n_theta, dtype='float32')
    probe_pos = np.array(probe_pos).astype(float)
    pos_int = np.round(probe_pos).astype(int)
    voxel_nm = np.array([psize_cm] * 3) * 1e7
    h = orc.get_kernel(voxel_nm[-1], 1240. / energy_ev, voxel_nm, P)
    cfg = dict(this_obj_size=[Y, X, Z], probe_size=P, energy_ev=energy_ev, psize_cm=psize_cm, h=h,
               free_prop_cm=free_prop_cm, unknown_type='delta_beta', optimize_all_probe_pos=optimize_all_probe_pos)
    pr, pi = orc.gaussian_probe(P, 1, probe_mag_sigma, probe_phase_max, probe_phase_sigma)
    pr = torch.tensor(pr[None], dtype=dtype)
    pi = torch.tensor(pi[None], dtype=dtype)
    luts = [orc.rotation_coords([Y, X, Z], th) for th in theta_ls]
    luts_rev = [orc.rotation_coords([Y, X, Z], -th) for th in theta_ls]      # util.py:505-512: the '_'-prefixed tables
    ppc = torch.tensor(np.tile(probe_pos - pos_int, [n_theta, 1, 1]), dtype=dtype)
    obj = torch.tensor(np.stack(initial_guess, -1), dtype=dtype)
    mask_t = None if mask is None else torch.tensor(mask, dtype=dtype).unsqueeze(-1)
    m = torch.zeros_like(obj)
    v = torch.zeros_like(obj)
    mp = torch.zeros((1, P[0], P[1], 2), dtype=dtype)
    vp = torch.zeros_like(mp)
    mq = torch.zeros_like(ppc)
    vq = torch.zeros_like(ppc)
    hist = []
    weight_l1 = None
    want = ['obj'] + (['probe_real', 'probe_imag'] if optimize_probe else []) + \
           (['probe_pos_correction'] if optimize_all_probe_pos else [])
    for i_epoch in range(n_epochs):
        chunks = orc.epoch_schedule(i_epoch, n_theta, len(pos_int), minibatch_size, 1, update_scheme=update_scheme)
        n_batch = len(chunks)
        i_opt = 0
        init_grads = True
        current_i_theta = -1
        obj_rot = None
        for i_batch in range(n_batch):
            i_theta, ind = orc.rank_slice_of_chunk(chunks, i_batch, 0, minibatch_size, 1)
            if rotate_out_of_loop and i_theta != current_i_theta:
                with torch.no_grad():
                    obj_rot = orc.apply_rotation(obj, luts[i_theta])       # ptychography.py:918-943, once per angle
            current_i_theta = i_theta
            last_of_theta = i_batch == n_batch - 1 or chunks[i_batch + 1][0, 0] != i_theta
            reg = dict(alpha_d=alpha_d, alpha_b=alpha_b, gamma=gamma)
            if reweighted_l1:
                if i_batch % 10 == 0:
                    weight_l1 = obj.max() / (obj.abs() + 1e-4 * obj.mean())
                reg['weight_l1'] = weight_l1
            loss, _, _, grads = orc.ptycho_loss_and_grads(
                obj_rot if rotate_out_of_loop else obj, pr, pi, pos_int[ind],
                torch.tensor(data_abs[i_theta, ind], dtype=dtype), None if rotate_out_of_loop else luts[i_theta], cfg, reg=reg,
                probe_pos_correction=ppc if optimize_all_probe_pos else None, this_i_theta=i_theta, this_ind_batch=ind,
                want=tuple(want))
            if init_grads:
                g_obj = torch.zeros_like(obj)
                g_probe = torch.zeros_like(mp)
                g_ppc = torch.zeros_like(ppc)
                init_grads = False
            g_obj = g_obj + grads['obj']
            if rotate_out_of_loop:
                g_obj = orc.apply_rotation(g_obj, luts_rev[i_theta])        # ptychography.py:1067-1078, every minibatch
            if optimize_probe:
                g_probe = g_probe + torch.stack([grads['probe_real'], grads['probe_imag']], -1)
            if optimize_all_probe_pos:
                g_ppc = g_ppc + grads['probe_pos_correction']
            if update_scheme == 'per angle' and not last_of_theta:
                continue
            init_grads = True
            obj, m, v = orc.adam_step(obj, g_obj, m, v, i_opt, step_size=learning_rate)
            if non_negativity:
                obj = obj.clamp(min=0)
            if optimize_probe:
                pstk, mp, vp = orc.adam_step(torch.stack([pr, pi], -1), g_probe, mp, vp, i_opt, step_size=probe_learning_rate)
                pr, pi = pstk[..., 0], pstk[..., 1]
            if optimize_all_probe_pos:
                ppc, mq, vq = orc.adam_step(ppc, g_ppc, mq, vq, i_opt, step_size=all_probe_pos_learning_rate)
                ppc = ppc - ppc.mean(dim=(0, 1))
            if mask_t is not None:
                obj = obj * mask_t
            hist.append((i_epoch, i_batch, float(loss)))
            if last_of_theta:
                i_opt += 1
    return np.array(hist), obj.numpy(), pr.numpy(), pi.numpy(), ppc.numpy()

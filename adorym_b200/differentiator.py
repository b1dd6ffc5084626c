"""adorym/differentiator.py:21-42 with the PyTorch-backend semantics of wrappers.py:300-331."""
import torch


class Differentiator(object):

    def __init__(self):
        self.loss_object = None
        self.opt_args_ls = []
        self.loss_args = {}

    def create_loss_node(self, loss, opt_args_ls=None):
        """:param loss: the loss function handle; :param opt_args_ls: indices (in the loss function's argument
        list) of the arguments to differentiate."""
        self.loss_object = loss
        self.opt_args_ls = opt_args_ls

    def get_gradients(self, **kwargs):
        """Returns the tuple of gradients ordered as opt_args_ls (wrappers.py:317-322)."""
        l = self.loss_object(**kwargs)
        nodes = [v for i, v in enumerate(kwargs.values()) if i in self.opt_args_ls]
        return torch.autograd.grad(l, nodes, retain_graph=False, create_graph=False, allow_unused=False)

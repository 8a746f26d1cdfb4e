"""Weight initialisation / device placement with the reference's semantics.

Mirrors model/networks.py:67-118 (`init_weights`, `init_net`): the net is moved to CUDA first, then every
Conv weight is drawn from N(0, init_gain^2) and every bias zeroed, in module registration order -- so the same
seed consumes the same RNG streams as the reference and yields the same parameters.
"""
import torch
import torch.nn as nn
from torch.nn import init


def init_weights(net, init_type='normal', init_gain=0.02):
    if init_type != 'normal':
        raise NotImplementedError('initialization method [%s] is not implemented' % init_type)

    def fn(m):
        # in-place on the Parameter itself (under no_grad), not on `.data` as the reference does: same values and RNG
        # consumption, but the tensor version is bumped, which is what keys the engine's packed-weight cache
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            with torch.no_grad():
                init.normal_(m.weight, 0.0, init_gain)
                if m.bias is not None:
                    init.constant_(m.bias, 0.0)

    print('initialize network with %s' % init_type)
    net.apply(fn)


def init_net(net, init_type='normal', init_gain=0.02):
    net.cuda()  # unconditional in the reference as well (model/networks.py:116)
    init_weights(net, init_type, init_gain=init_gain)
    return net

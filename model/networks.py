from wacv23_tsnet_b200.model.networks import init_net, init_weights  # noqa: F401

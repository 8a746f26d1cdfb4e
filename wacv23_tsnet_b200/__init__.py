"""B200-native forward hot path of TS-Net (WACV'23): hand-written sm_100a kernels behind a C ABI,
exposed through the reference's own `TSNet(...)` class surface (wacv23_tsnet_b200.model)."""
__version__ = "0.1.0"

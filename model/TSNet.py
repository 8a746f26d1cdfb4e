from wacv23_tsnet_b200.model.TSNet import *  # noqa: F401,F403
from wacv23_tsnet_b200.model.TSNet import TSNet  # noqa: F401

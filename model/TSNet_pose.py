from wacv23_tsnet_b200.model.TSNet_pose import TSNet  # noqa: F401

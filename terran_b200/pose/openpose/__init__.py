from terran_b200.pose.openpose.wrapper import OpenPose  # noqa

"""poselib_b200 — B200-native LO-RANSAC / minimal-solver engine behind PoseLib's call surface."""
__version__ = "0.1.0"

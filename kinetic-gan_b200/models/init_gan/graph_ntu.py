"""NTU RGB+D 25-joint skeleton, 25 -> 11 -> 5 -> 1 (reference: models/init_gan/graph_ntu.py:5-21)."""
from .skeleton import SkeletonGraph


class graph_ntu(SkeletonGraph):
    def __init__(self, max_hop=1, dilation=1):
        super().__init__("ntu", max_hop, dilation)

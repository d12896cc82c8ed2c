"""Human3.6M 16-joint skeleton, 16 -> 7 -> 2 -> 1, centre = thorax (reference: models/init_gan/graph_h36m.py:5-21,29)."""
from .skeleton import SkeletonGraph


class Graph_h36m(SkeletonGraph):
    def __init__(self, max_hop=1, dilation=1):
        super().__init__("h36m", max_hop, dilation)

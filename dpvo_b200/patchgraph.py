"""Device-resident patch graph: fixed-capacity edge arrays + per-edge recurrent state (csrc/pgraph.cu).

The reference's patch graph (dpvo/patchgraph.py:26-35 ii / jj / kk / net, dpvo/dpvo.py:215-238 append_factors /
remove_factors, :266-310 keyframe) grows and shrinks every frame with torch.cat and boolean-mask indexing, which costs
host synchronisations and changes tensor shapes -- a captured CUDA graph of update() cannot follow it.  Here the arrays
have a fixed capacity; a slot holds an active edge or a parked dummy edge that points at a reserved frame / patch and so
lives in groups of its own in every grouping of the update operator.  Removal parks slots in place, append fills parked
slots in index order and zeroes the state rows of the new edges; all decisions read device scalars.
"""
import torch

from . import extensions
from .net import DIM


class DevicePatchGraph:
    def __init__(self, capacity, M, dummy_frame, device="cuda"):
        """capacity: edge slots (rounded up to a multiple of 128); dummy_frame: a frame slot no real edge ever uses (its
        pose / intrinsics / patches must be valid data, e.g. the last slot of the pose buffer)."""
        self.cap = (int(capacity) + 127) // 128 * 128
        self.M = int(M)
        self.dummy_frame = int(dummy_frame)
        self.dummy_patch = self.dummy_frame * self.M
        dev = torch.device(device)
        self.ii = torch.full((self.cap,), self.dummy_frame, dtype=torch.long, device=dev)
        self.jj = torch.full((self.cap,), self.dummy_frame, dtype=torch.long, device=dev)
        self.kk = torch.full((self.cap,), self.dummy_patch, dtype=torch.long, device=dev)
        self.active = torch.zeros(self.cap, dtype=torch.uint8, device=dev)
        self.n_active = torch.zeros(1, dtype=torch.int32, device=dev)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self.net = torch.zeros(1, self.cap, DIM, device=dev)                  # the recurrent state `net` (dpvo.py:53), fp32

    # -------------------------------------------------------------------------------- dpvo.py:215-222
    def append(self, ii, jj, kk, enable=None):
        """add edges (ii = source frame, jj = target frame, kk = patch); their state rows start at zero.
        Returns the slot of every new edge (int32, -1 = store full, `overflow` is then set)."""
        return extensions()[3].pgraph_append(self.ii, self.jj, self.kk, self.active, ii, jj, kk, enable, self.net,
                                             self.n_active, self.overflow)

    def append_frame_edges(self, n_dev, lifetime, enable=None):
        """the forward + backward edges of the newest frame n - 1 (dpvo.py:457-459 with :362-375), n on the device"""
        ii, jj, kk = extensions()[3].pgraph_new_edges(n_dev, self.M, lifetime)
        return self.append(ii, jj, kk, enable)

    # -------------------------------------------------------------------------------- dpvo.py:224-238, 300-306
    def remove_old(self, n_dev, window, enable=None):
        """park the edges whose patch left the removal window: ix[kk] < n - window"""
        extensions()[3].pgraph_remove(self.ii, self.jj, self.kk, self.active, 0, n_dev, int(window), enable,
                                      self.dummy_frame, self.dummy_patch, self.M, self.n_active)

    # -------------------------------------------------------------------------------- dpvo.py:279-286
    def remove_frame(self, k_dev, enable=None):
        """keyframe removal: park the edges that touch frame k and renumber frames / patches above it"""
        extensions()[3].pgraph_remove(self.ii, self.jj, self.kk, self.active, 1, k_dev, 0, enable,
                                      self.dummy_frame, self.dummy_patch, self.M, self.n_active)

    # -------------------------------------------------------------------------------- host-side views (synchronise)
    def edges(self):
        m = self.active.bool()
        return self.ii[m], self.jj[m], self.kk[m]

    def net_rows(self):
        """[cap, 384] view of the recurrent state"""
        return self.net[0]

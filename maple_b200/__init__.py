"""maple_b200: B200-native (sm_100a) SPR-likelihood kernels behind MAPLE's own interfaces.

Only what the hot path needs lives here: the packed genome-list format (genome_list), the model
state the kernels read (model), the ctypes binding of the C ABI (capi), and the host-side mirror
of the reference's functions for this path (engine).
"""
from .genome_list import PackedLists, pack_lists, unpack_list, lists_equal  # noqa: F401
from .model import MapleModel  # noqa: F401

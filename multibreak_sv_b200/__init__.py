"""multibreak_sv_b200 — B200-native (CUDA sm_100a) drop-in for the MCMC sampler of raphael-group/multibreak-sv.

Only the sampler hot path lives here: the C-ABI library `libmbsv_b200.so` (csrc/), its ctypes
binding (`_abi`), and the host-side mirror of the reference's `MultiBreakSV` driver.
"""
from ._abi import (MODE_FAST, MODE_REPLAY_EXACT, DeviceProblem, MbsvError, PackedProblem, RunOutput, load_library,
                   reference_burnin)

__all__ = ["MODE_FAST", "MODE_REPLAY_EXACT", "DeviceProblem", "MbsvError", "PackedProblem", "RunOutput",
           "load_library", "reference_burnin"]

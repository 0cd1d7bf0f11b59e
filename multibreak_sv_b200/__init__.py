"""multibreak_sv_b200 — B200-native (CUDA sm_100a) drop-in for the MCMC sampler of raphael-group/multibreak-sv.

Only the sampler hot path lives here: the C-ABI library `libmbsv_b200.so` (csrc/), its ctypes
binding (`_abi`), and the host-side mirror of the reference's `MultiBreakSV` driver.
"""
import os as _os

# A run overlaps one kernel per state-size class (up to 19 streams); with CUDA's default of 8 hardware queues the 9th
# stream would silently wait for the 1st.  CUDA reads the variable when the context is created, so the HOST sets it before
# its first CUDA call (the library itself never touches the environment; include/mbsv_b200.h "Host setup").
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from ._abi import (MODE_FAST, MODE_GIBBS, MODE_REPLAY_EXACT, Comm, DeviceProblem, MbsvError, PackedProblem, RunOutput,
                   load_library, reference_burnin, run_multi, trim)

__all__ = ["MODE_FAST", "MODE_GIBBS", "MODE_REPLAY_EXACT", "Comm", "DeviceProblem", "MbsvError", "PackedProblem", "RunOutput",
           "load_library", "reference_burnin", "run_multi", "trim"]

"""B200-native guidance + alignment hot path of FollowMyHold."""
__version__ = "0.1.0"

import os as _os

# One evaluation forks onto three library side streams per lane, the host API adds an upload and a
# download stream: more concurrent streams than the default 8 hardware work queues.  Streams that alias
# onto one queue serialise behind each other (a 10 ms upload then stalls a compute chain).  The variable
# is read when the CUDA context is created, so it has to be set before the first CUDA call.
_os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

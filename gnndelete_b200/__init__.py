"""B200-native implementation of GNNDelete's unlearning hot path.

Host side: Python/PyTorch (device memory, streams, torch.distributed).
Arithmetic: hand-written sm_100a CUDA kernels in ``csrc/`` reached through the
C ABI declared in ``include/gnndelete_b200.h`` (``libgnndelete_b200.so``).
There is no CPU fallback: importing the kernel bindings without the built
library raises.
"""
__version__ = '0.1.0'

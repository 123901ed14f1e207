"""B200-native (sm_100a) photometric view-synthesis loss path of unsupervised depth / optical-flow / ego-motion
training: hand-written CUDA kernels behind a C-ABI (``include/ugl.h``), exposed with the reference's own call
signatures.  See DESIGN.md and INTEGRATION.md."""
__version__ = "0.1.0"

"""vins-mobile_b200: B200-native VIO hot path (front-end tracker + sliding-window optimiser).

The product is the C-ABI shared library `libvio_b200.so` (include/vio_b200.h) built from csrc/*.cu for
sm_100a; this package only holds the ctypes binding (`api`), the ABI struct (`abi`) and the synthetic
stream generator used by tests and bench (`synth`).  Import with
`importlib.import_module("vins-mobile_b200")` (the directory name is not a Python identifier)."""

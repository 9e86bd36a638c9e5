"""B200-native PIC-DSMC particle loop: CUDA kernels (csrc/) behind the C ABI of include/picgpu.h,
a ctypes binding with the reference's class names (picgpu.py) and the C++ facade (host/).
The package name contains hyphens; import it with importlib.import_module(<dir name>)."""

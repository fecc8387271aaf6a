"""Drop-in ``sfast`` package (B200 build).

Only the surface the reference declares stable is provided -- ``sfast.compilers``
(/root/reference/README.md:133) -- backed by hand-written sm_100a kernels behind a C ABI
(``sfast_b200/libsfb200.so``) instead of TorchScript passes + cuDNN/cuBLASLt/CUTLASS/Triton.
"""
__version__ = "1.0.5+b200"

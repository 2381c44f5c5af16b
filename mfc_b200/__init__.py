"""mfc_b200 -- B200-native drop-in for MFC's 5-equation RHS hot path.

Host-side mirror of the reference interface for that path only:
``params`` (case.py -> parameter struct), ``pre_process`` (grid + patch ICs), ``decomp`` (block
decomposition), ``lib`` (ctypes binding of the C-ABI CUDA library ``libmfc5.so``) and ``driver``
(``s_tvd_rk`` time loop, ``D/*.dat`` output).  The CUDA kernels live in ``csrc/``.
"""

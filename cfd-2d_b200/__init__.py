"""cfd-2d_b200: B200-native (sm_100a) implementation of ONE path of zhrv/cfd-2d -- the explicit
finite-volume residual + RK2 update on unstructured triangle meshes (``FVM_TVD``,
reference ``src/methods/fvm_tvd.cpp``).

Layout
  csrc/      hand-written CUDA kernels + the C-ABI (include/cfd2d_fvm.h) -> libcfd2d_b200.so
  host/      C++ glue: ``FVM_TVD_CUDA : public Method`` compiled against the reference headers
  mesh.py    nodes/triangles -> flat SoA (restates the reference UNV reader), UNV writer
  task.py    task.xml schema (reader/writer)
  cases.py   the synthetic cases of SURVEY.md section 8(d)
  fvm.py     ctypes binding of the C-ABI + ``FVM_TVD`` host mirror (init/run/done)
  decomp.py  METIS partition + owned/halo renumbering (restates reference Decomp)

There is no CPU fallback: everything numerical goes through libcfd2d_b200.so and fails loudly
if it is missing.
"""
__version__ = "0.1.0"

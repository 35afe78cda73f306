import ctypes, os, sys
lib = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "..", "mrcpp_b200", "lib", sys.argv[1] if len(sys.argv) > 1 else "libmb_test.so"))
for f in ("mrx_bench_dmma_tflops", "mrx_bench_dfma_tflops"):
    getattr(lib, f).restype = ctypes.c_double
    getattr(lib, f).argtypes = [ctypes.c_int]
lib.mrx_bench_hbm_gbs.restype = ctypes.c_double
lib.mrx_bench_hbm_gbs.argtypes = [ctypes.c_longlong, ctypes.c_int]
print("DMMA TFLOP/s", lib.mrx_bench_dmma_tflops(20000))
print("DFMA TFLOP/s", lib.mrx_bench_dfma_tflops(20000))
print("HBM copy GB/s", lib.mrx_bench_hbm_gbs(2 << 30, 10))

"""Developer check: attention kernels at the bench shapes vs the fp64 reference (errors printed, nothing asserted)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_kernels_gpu import attn_bench_shape_errors

for kind in ("enc", "dec", "cross"):
    errs = attn_bench_shape_errors(kind)
    print(kind, " ".join("%s: max %.2e rel %.2e (ref %.2e)" % (k, *v) for k, v in errs.items()), flush=True)

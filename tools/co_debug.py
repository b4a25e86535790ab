"""Reproduce the co-scheduled numpy-path failure at 4096 (B=1) -- run under compute-sanitizer."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["ADRT_B200_COSCHED"] = "1"
import adrt_b200 as adrt
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
x = np.random.default_rng(0).standard_normal((1, n, n)).astype(np.float32)
y = adrt.adrt(x)
print("adrt ok", float(y.sum()))
z = adrt.bdrt(y)
print("bdrt ok", float(z[0, 0, :4, :4].sum()))

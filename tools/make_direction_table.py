"""Generate globalsensitivity.jl_b200/data/joe_kuo_21201.bin from SciPy's copy of the
public Joe-Kuo `new-joe-kuo-6.21201` table (the same table Sobol.jl ships).

Layout: little-endian uint32, 21201 records of 19 words: [poly, m_1 .. m_18]
  poly   = primitive polynomial INCLUDING leading and trailing 1 bits (3 = x+1)
  m_i    = initial direction integers (only the first deg(poly) are meaningful)
Record 0 is the van der Corput dimension (poly=1, all m=1 handled by the consumer).

Run:  python tools/make_direction_table.py
"""
import os, sys
import numpy as np
import scipy.stats

src = os.path.join(os.path.dirname(scipy.stats.__file__), "_sobol_direction_numbers.npz")
z = np.load(src)
poly = z["poly"].astype(np.uint32)
vinit = z["vinit"].astype(np.uint32)
assert poly.shape == (21201,) and vinit.shape == (21201, 18)
rec = np.concatenate([poly[:, None], vinit], axis=1).astype("<u4")
out = os.path.join(os.path.dirname(__file__), "..", "globalsensitivity.jl_b200", "data", "joe_kuo_21201.bin")
rec.tofile(out)
print("wrote", os.path.abspath(out), rec.nbytes, "bytes")

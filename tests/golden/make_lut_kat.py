"""Generate tests/golden/lut_kat_d128_m32.npz from the reference's literal LUT known-answer test.

Source: /root/reference/Chameleon/retrieval_accelerator/LUT_construction_PEs/
        LUT_construction_PE_D128_M32/src/host.cpp
  :44-46   product_quantizer[i] = i % 256 over the flat (M, 256, D/M) array
  :58-65   query_vec_data[128]   (literal)
  :74-81   center_vec_data[128]  (literal)
  :92-109  software LUT: LUT[row][m] = sum_c (diff[m*D/M+c] - pq[m][row][c])^2
  :176-191 pass rule: |hw - sw| <= 0.01 and 0.99 <= hw/sw <= 1.01

The literal vectors are parsed out of host.cpp (this script only runs in the build container, where
/root/reference exists); the expected table is computed in exact integer arithmetic, independently of
oracle/.  Every input is an integer and every partial sum is < 2^24, so the fp32 result is exact and
the golden table pins the oracle bit for bit.
"""
import os
import re

import numpy as np

SRC = ("/root/reference/Chameleon/retrieval_accelerator/LUT_construction_PEs/"
       "LUT_construction_PE_D128_M32/src/host.cpp")
D, M, KSUB = 128, 32, 256


def literal(text, name):
    m = re.search(name + r"\[D\]\s*=\s*\{([^}]*)\}", text, re.S)
    vals = [int(v) for v in m.group(1).replace("\n", " ").split(",")]
    assert len(vals) == D, (name, len(vals))
    return np.array(vals, dtype=np.int64)


def main():
    text = open(SRC).read()
    q = literal(text, "query_vec_data")
    c = literal(text, "center_vec_data")
    dsub = D // M
    pq = (np.arange(M * KSUB * dsub, dtype=np.int64) % KSUB).reshape(M, KSUB, dsub)
    res = (q - c).reshape(M, 1, dsub)
    T = ((res - pq) ** 2).sum(axis=2)          # (M, 256), exact integers
    assert T.max() < 2 ** 24
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lut_kat_d128_m32.npz")
    np.savez_compressed(out, query=q.astype(np.float32), centroid=c.astype(np.float32),
                        pq=pq.astype(np.float32), lut=T.astype(np.float32))
    print("wrote", out, "lut[0,:4] =", T[0, :4], "max", T.max())


if __name__ == "__main__":
    main()

"""numpy check of the mixed-radix Stockham indexing used by csrc/kernels_fft.cu"""
import numpy as np

def dft_small(x, inv):
    r = len(x); sgn = 1.0 if inv else -1.0
    return [sum(x[k] * np.exp(sgn * 2j * np.pi * j * k / r) for k in range(r)) for j in range(r)]

def stockham(x, radices, inv=False):
    K = len(x); src = np.array(x, complex); dst = np.zeros(K, complex)
    sgn = 1.0 if inv else -1.0
    W = np.exp(sgn * 2j * np.pi * np.arange(K) / K)
    n, s = K, 1
    for r in radices:
        m = n // r
        for idx in range(K // r):
            pp, q = divmod(idx, s)
            a = [src[q + s * (pp + k * m)] for k in range(r)]
            b = dft_small(a, inv)
            for j in range(r):
                dst[q + s * (r * pp + j)] = b[j] * W[(j * pp * (K // n)) % K]
        src, dst = dst, src
        n, s = m, s * r
    return src

for K, rad in ((48, [4, 4, 3]), (64, [4, 4, 4]), (32, [4, 4, 2]), (36, [4, 3, 3]), (96, [4, 4, 3, 2])):
    x = np.random.default_rng(0).normal(size=K) + 1j * np.random.default_rng(1).normal(size=K)
    print(K, rad, np.abs(stockham(x, rad) - np.fft.fft(x)).max(), np.abs(stockham(x, rad, True) - np.fft.ifft(x) * K).max())

"""Seeded synthetic inputs shared by the parity tests (shapes the oracle finishes in seconds)."""
import numpy as np


def genotype(seed, n, p, n_classes=2, dtype=np.float64):
    rs = np.random.RandomState(seed)
    x = rs.randint(0, 3, (n, p))
    y = rs.randint(0, n_classes, n)
    risk = (x[:, 1 % p] == 1) & (x[:, 3 % p] == 1)
    y[risk] = 1
    return x.astype(dtype), y.astype(np.int64)


def gaussian(seed, n, p, n_classes=2, dtype=np.float64):
    rs = np.random.RandomState(seed)
    y = rs.randint(0, n_classes, n)
    x = rs.standard_normal((n, p))
    x[:, 0] += 1.5 * y
    x[:, 2 % p] -= 1.0 * y
    return x.astype(dtype), y.astype(np.int64)


def mixed(seed, n, p, n_classes=3, dtype=np.float64):
    rs = np.random.RandomState(seed)
    y = rs.randint(0, n_classes, n)
    x = np.empty((n, p))
    h = p // 2
    x[:, :h] = rs.randint(0, 3, (n, h))
    x[:, h:] = rs.standard_normal((n, p - h))
    x[:, 0] = (y + rs.randint(0, 2, n)) % 3
    x[:, h] += 1.2 * y
    x[:, p - 1] = 7.0          # constant column
    if p > 6:
        x[:, 4] = rs.randint(0, 7, n) * 0.5      # 7-valued discrete column
    return x.astype(dtype), y.astype(np.int64)


def epistatic_genotypes(seed, n, p):
    """Generator of SURVEY.md section 8(d) C3/C5 (after benchmarking/BenchmarkingRelief2.ipynb cell 9):
    int8 0/1/2 genotypes, cases where SNP 25 and SNP 75 are both heterozygous, then random
    controls promoted until half the samples are cases."""
    rs = np.random.RandomState(seed)
    x = rs.randint(0, 3, (n, p)).astype(np.int8)
    y = np.zeros(n, np.int64)
    a, b = 25 % p, 75 % p
    y[(x[:, a] == 1) & (x[:, b] == 1)] = 1
    need = n // 2 - int(y.sum())
    if need > 0:
        ctrl = np.flatnonzero(y == 0)
        y[rs.choice(ctrl, need, replace=False)] = 1
    return x, y

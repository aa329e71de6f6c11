"""CPU: the algebra of csrc/mixed_radix.cu (orders 3*2^k and 9*2^k) restated in plain Python integers and checked against the
transform by definition with GF_Root(N) (Slow_NTT, ntt.cpp:451-483): the Cooley-Tukey split n = n1 + r*n2 / k = k2 + M*k1, the
CRT split of the twiddle w_N^e into GF_Root(r)^(a*e mod r) * GF_Root(M)^(b*e mod M), the order-3 codelet in the form of the
reference's NTT3 (ntt.cpp:26-46), the 3 x 3 four-step order-9 codelet (ntt.cpp:114-146) with its transposed output, and the
inverse direction.  The GPU kernel itself is compared with the oracle in tests/test_gpu_parity.py."""
import random

import pytest

P = 0xFFF00001


def root(n):
    return pow(19, (P - 1) // n, P)                       # GF_Root, GF(p).cpp:267-276


def by_definition(x, N, inverse):
    w = root(N)
    if inverse:
        w = pow(w, P - 2, P)
    return [sum(x[n] * pow(w, n * k, P) for n in range(N)) % P for k in range(N)]


def dft3(f, z3):
    x1, x2, inv2 = z3, z3 * z3 % P, (P + 1) // 2
    c1, c2 = (x1 + x2) * inv2 % P, (x1 - x2) * inv2 % P
    s, d = (f[1] + f[2]) % P, (f[1] - f[2]) % P
    u, v = (f[0] + s * c1) % P, d * c2 % P
    return [(f[0] + s) % P, (u + v) % P, (u - v) % P]


def mixed(x, r, M, inverse):
    Y = [by_definition([x[n1 + r * n2] for n2 in range(M)], M, inverse) for n1 in range(r)]     # step 1: the power-of-two transforms
    zr, z3, gM = root(r), root(3), (root(M) if M > 1 else 1)
    if inverse:
        zr, z3 = pow(zr, P - 2, P), pow(z3, P - 2, P)
    a = [a for a in range(r) if a * (M % r) % r == 1 % r][0]
    b = pow(r, -1, M) if M > 1 else 0
    out = [0] * (r * M)
    for k2 in range(M):
        f = []
        for n1 in range(r):
            e = n1 * k2
            er, em = a * (e % r) % r, (b * (e % M) % M if M > 1 else 0)
            if inverse:
                em = (M - em) % M                          # the power table is forward; zr already has the direction
            f.append(Y[n1][k2] * pow(gM, em, P) % P * pow(zr, er, P) % P)
        if r == 3:
            g, pos = dft3(f, z3), [0, 1, 2]
        else:
            g = f[:]
            for c in range(3):
                g[c], g[3 + c], g[6 + c] = dft3([g[c], g[3 + c], g[6 + c]], z3)
            g[4] = g[4] * zr % P; g[5] = g[5] * pow(zr, 2, P) % P; g[7] = g[7] * pow(zr, 2, P) % P; g[8] = g[8] * pow(zr, 4, P) % P
            for c in range(3):
                g[3 * c], g[3 * c + 1], g[3 * c + 2] = dft3([g[3 * c], g[3 * c + 1], g[3 * c + 2]], z3)
            pos = [(k1 % 3) * 3 + k1 // 3 for k1 in range(9)]
        for k1 in range(r):
            out[k2 + M * k1] = g[pos[k1]]
    return out


@pytest.mark.parametrize("r,M", [(3, 1), (9, 1), (3, 2), (3, 8), (9, 4), (9, 16), (3, 32)])
def test_mixed_radix_decomposition_equals_the_definition(r, M):
    random.seed(100 * r + M)
    x = [random.randrange(P) for _ in range(r * M)]
    for inverse in (False, True):
        assert mixed(x, r, M, inverse) == by_definition(x, r * M, inverse)

"""TEST INFRASTRUCTURE ONLY: erasure decoding BY DEFINITION for small orders.  The reference has no decoder (roadmap
README.md:173; algorithm sketches README.md:88-119, RS.md:42-79), so parity for this row is UNPINNED by reference code:
this file is the definition -- the unique polynomial of degree < N through any N known symbols c[m] = f(rho^m),
rho = GF_Root(2N) (GF(p).cpp:267-276), evaluated at the erased points by Lagrange's formula in Python integers."""
P = 0xFFF00001


def recover(code, erased):
    """code: list of 2N rows (lists of ints; erased rows ignored); erased: positions.  Returns {position: row}."""
    n2 = len(code)
    n = n2 // 2
    rho = pow(19, (P - 1) // n2, P)
    es = set(erased)
    keep = [m for m in range(n2) if m not in es][:n]
    assert len(keep) == n, "more than N erasures"
    xs = [pow(rho, m, P) for m in keep]
    dinv = []
    for a, xa in enumerate(xs):
        d = 1
        for b, xb in enumerate(xs):
            if a != b:
                d = d * (xa - xb) % P
        dinv.append(pow(d, P - 2, P))
    out = {}
    for e in erased:
        xe = pow(rho, e, P)
        diffs = [(xe - xb) % P for xb in xs]
        acc = [0] * len(code[keep[0]])
        for a in range(n):
            num = 1
            for b, df in enumerate(diffs):
                if a != b:
                    num = num * df % P
            coef = num * dinv[a] % P
            acc = [(s + coef * v) % P for s, v in zip(acc, code[keep[a]])]
        out[e] = acc
    return out

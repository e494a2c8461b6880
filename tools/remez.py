"""Weighted Remez exchange in mpmath (generator-side only; not shipped in the product path).

minimise  max_{z in [a,b]} | w(z) * (P(z) - g(z)) |   over polynomials P of degree d.
Used by tools/gen_constants.py to fit the sin/cos kernels and the Psi-Hyperbasis correction terms.
"""
import mpmath as mp


def _solve(nodes, g, w, d):
    # unknowns: c_0..c_d, E ;  w(z_i) * (P(z_i) - g(z_i)) = (-1)^i E
    n = d + 2
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for i, z in enumerate(nodes):
        wz = w(z)
        for j in range(d + 1):
            A[i, j] = wz * z ** j
        A[i, d + 1] = -((-1) ** i)
        b[i] = wz * g(z)
    sol = mp.lu_solve(A, b)
    return [sol[j] for j in range(d + 1)], sol[d + 1]


def polyval(c, z):
    acc = mp.mpf(0)
    for cj in reversed(c):
        acc = acc * z + cj
    return acc


def remez(g, w, d, a, b, iters=40, grid=4000, verbose=False):
    a = mp.mpf(a)
    b = mp.mpf(b)
    n = d + 2
    nodes = [a + (b - a) * (1 - mp.cos(mp.pi * (2 * i + 1) / (2 * n))) / 2 for i in range(n)]
    coeffs, E = None, None
    for it in range(iters):
        coeffs, E = _solve(nodes, g, w, d)
        err = lambda z: w(z) * (polyval(coeffs, z) - g(z))
        # dense scan for local extrema of the error function
        zs = [a + (b - a) * (1 - mp.cos(mp.pi * i / grid)) / 2 for i in range(grid + 1)]
        es = [err(z) for z in zs]
        ext = []
        for i in range(grid + 1):
            left = es[i - 1] if i > 0 else None
            right = es[i + 1] if i < grid else None
            e = es[i]
            is_max = (left is None or abs(e) >= abs(left)) and (right is None or abs(e) >= abs(right))
            if is_max and e != 0:
                ext.append((zs[i], e))
        # keep an alternating subsequence of the largest extrema
        alt = []
        for z, e in ext:
            if alt and (e > 0) == (alt[-1][1] > 0):
                if abs(e) > abs(alt[-1][1]):
                    alt[-1] = (z, e)
            else:
                alt.append((z, e))
        while len(alt) > n:
            # drop the smaller of the two ends (keeps alternation)
            if abs(alt[0][1]) < abs(alt[-1][1]):
                alt.pop(0)
            else:
                alt.pop()
        if len(alt) < n:
            break
        new_nodes = [z for z, _ in alt]
        emax = max(abs(e) for _, e in alt)
        emin = min(abs(e) for _, e in alt)
        if verbose:
            print(f"  remez it={it} E={mp.nstr(E, 6)} emax={mp.nstr(emax, 6)} ratio={mp.nstr(emax / emin, 6)}")
        nodes = new_nodes
        if emax / emin < mp.mpf("1.0005"):
            break
    coeffs, E = _solve(nodes, g, w, d)
    return coeffs, abs(E)

"""Host-side secp256k1 on Python ints: control-path arithmetic only (window base points, key
recovery).  Mirrors the conventions of lib/Curve64.pb that are observable by the GPU path:
m_ADDPTX64 doubles when the x coordinates are equal (Curve64.pb:2560), points are affine."""
P = 2**256 - 2**32 - 977
N = 0xFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFFEBAAEDCE6AF48A03BBFD25E8CD0364141
G = (0x79BE667EF9DCBBAC55A06295CE870B07029BFCDB2DCE28D959F2815B16F81798,
     0x483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8)


def add(a, b):
    """Affine add; None is the point at infinity."""
    if a is None:
        return b
    if b is None:
        return a
    if a[0] == b[0]:
        if (a[1] + b[1]) % P == 0:
            return None
        lam = 3 * a[0] * a[0] * pow(2 * a[1], -1, P) % P
    else:
        lam = (a[1] - b[1]) * pow(a[0] - b[0], -1, P) % P
    x = (lam * lam - a[0] - b[0]) % P
    return (x, (lam * (a[0] - x) - a[1]) % P)


def neg(a):
    return None if a is None else (a[0], (-a[1]) % P)


def mul(a, k):
    k %= N
    r = None
    while k:
        if k & 1:
            r = add(r, a)
        a = add(a, a)
        k >>= 1
    return r


def y_from_x(x, odd):
    y = pow((x * x * x + 7) % P, (P + 1) // 4, P)          # m_YfromX64, Curve64.pb:2656-2682
    if (y * y - x * x * x - 7) % P:
        raise ValueError("x is not on the curve")
    return y if (y & 1) == odd else P - y


def parse_pubkey(s):
    """128 hex | '04'+128 | '02'/'03'+64 (1_9_7File.pb:5006-5018)."""
    s = s.strip().lower()
    if s.startswith("0x"):
        s = s[2:]
    if len(s) == 130 and s.startswith("04"):
        s = s[2:]
    if len(s) == 128:
        return (int(s[:64], 16), int(s[64:], 16))
    if len(s) == 66 and s[:2] in ("02", "03"):
        x = int(s[2:], 16)
        return (x, y_from_x(x, int(s[:2], 16) & 1))
    raise ValueError("unsupported public key format")


def compress(pt):
    return ("03" if pt[1] & 1 else "02") + format(pt[0], "064x")


def parse_w(arg):
    """-w: values <= 32 mean Int(2^x) (truncation), otherwise a decimal count (1_9_7File.pb:1013-1021)."""
    v = float(arg)
    return int(2.0 ** v) if v <= 32 else int(arg)

"""Pin the CPU oracle against every known-answer vector of the reference's own self-test block
(lib/Curve64.pb:3067-3397).  These are the reference's golden vectors for field/curve arithmetic."""
import random
from tests import pyec

H = lambda s: int(s.replace(" ", ""), 16)
GX, GY = pyec.G
P = pyec.P


def test_negmod_gy(orc):                       # Curve64.pb:3088-3095
    assert orc.fe1("orc_fe_neg", GY) == H("b7c52588d95c3b9a a25b0403f1eef757 02e84bb7597aabe6 63b82f6f04ef2777")


def test_mulmod_small(orc):                    # Curve64.pb:3097-3109 (modulus argument is ignored: always p)
    assert orc.fe2("orc_fe_mul", 3, 4) == 12


def test_add_g_2g(orc):                        # Curve64.pb:3113-3132
    g2 = (H("c6047f9441ed7d6d3045406e95c07cd85c778e4b8cef3ca7abac09b95c709ee5"),
          H("1ae168fea63dc339a3c58419466ceaeef7f632653266d0e1236431a950cfe52a"))
    assert orc.pt_add((GX, GY), g2) == (H("f9308a019258c31049344f85f89d5229b531c845836f99b08601f113bce036f9"),
                                        H("388f7b0f632de8140fe337e62a37f3566500a99934c2231b6cb9fd7584b8e672"))


def test_squares(orc):                         # Curve64.pb:3135-3168
    kat = [("342119815c0f816f31f431a9fe98a6c76d11425ecaeaecf2d0ef6def197c56b0", "38f37014ce22fc29cf19f28a5ce4da091445536c3e2cff318ba07c2a3048f518"),
           ("3fdc2a05828a06c18e057a8d9549bdc3ff05ee69a352342ce382aafeaeb98ef9", "3d6c452d1c076d0425ac63c7783f563df3ec12324d0f16bf7c8335253ef4be33"),
           ("483ADA7726A3C4655DA4FBFC0E1108A8FD17B448A68554199C47D08FFB10D4B8", "4866d6a5ab41ab2c6bcc57ccd3735da5f16f80a548e5e20a44e4e9b8118c26f2")]
    for a, r in kat:
        assert orc.fe1("orc_fe_sqr", H(a)) == H(r)
        assert orc.fe2("orc_fe_mul", H(a), H(a)) == H(r)


A_ = H("3fdc2a05828a06c18e057a8d9549bdc3ff05ee69a352342ce382aafeaeb98ef9")
B_ = H("dfcad171d3196bdb20eaaf272f8f9bcc6b5a47d4fe53d3d874e703cd2566197e")


def test_raw_add_sub(orc):                     # Curve64.pb:3203-3228
    import ctypes as C
    r = orc.Fe(); a = orc.Fe.of(A_); b = orc.Fe.of(B_)
    carry = orc.lib().orc_raw_add(C.byref(r), C.byref(a), C.byref(b))
    assert (carry << 256) + r.int() == H("11fa6fb7755a3729caef029b4c4d959906a60363ea1a608055869aecbd41fa877")
    borrow = orc.lib().orc_raw_sub(C.byref(r), C.byref(a), C.byref(b))
    assert borrow == 1 and r.int() == H("60115893af709ae66d1acb6665ba21f793aba694a4fe60546e9ba7318953757b")


def test_addmod_submod(orc):                   # Curve64.pb:3230-3255
    assert orc.fe2("orc_fe_add", A_, B_) == H("1fa6fb7755a3729caef029b4c4d959906a60363ea1a608055869aeccd41fac48")
    assert orc.fe2("orc_fe_sub", A_, B_) == H("60115893af709ae66d1acb6665ba21f793aba694a4fe60546e9ba730895371aa")


def test_inverse_gx(orc):                      # Curve64.pb:3260-3267
    assert orc.fe1("orc_fe_inv", GX) == H("237afdf1d2938d86870aaeb8ad77626a67b8e794abfb076be61d003687ca9ef6")
    assert orc.fe1("orc_fe_inv", 0) == 0       # Curve64.pb:2475-2477


def test_double_g_2g(orc):                     # Curve64.pb:3270-3299
    g2 = orc.pt_dbl((GX, GY))
    assert g2 == (H("c6047f9441ed7d6d3045406e95c07cd85c778e4b8cef3ca7abac09b95c709ee5"),
                  H("1ae168fea63dc339a3c58419466ceaeef7f632653266d0e1236431a950cfe52a"))
    assert orc.pt_dbl(g2) == (H("e493dbf1c10d80f3581e4904930b1404cc6c13900ee0758474fa94abe8c4cd13"),
                              H("51ed993ea0d455b75642e2098ea51448d967ae33bfbdfe40cfe97bdc47739922"))


def test_add_g_10000_times(orc):               # Curve64.pb:3303-3320
    b = (GX, GY)
    for _ in range(10000):
        b = orc.pt_add((GX, GY), b)
    assert b == (H("db7432110ba814bfe6371ddfd03ba554b558548aa90e81b8e1421321656065a8"),
                 H("8236f24d965a900384b382e8d772d7e92dee2ce6c3cb33883ea627d54a5170c4"))


def test_ptmul(orc):                           # Curve64.pb:3371-3392
    k = H("3fdc2a05828a06c18e057a8d9549bdc3ff05ee69a352342ce382aafeaeb98ef9")
    assert orc.pt_mul((GX, GY), k) == (H("510f6efbef396a1985da989104a295063606319beafa4e1fd0ebd29ace19088f"),
                                       H("fcf1cb9e1a9c02fea09e983fe5fe8fb7ce74a80ed3b1783706e27bde4b2ede5e"))


def test_sample_problems(orc):                 # 1_9_7File.pb:189,191 and the puzzle keys of BASELINE.json
    pub = orc.parse_pubkey("e1e5e6f7b0b8d67604e3940c87bf06b814cedc486112b9956c68e3d78b1bd812"
                           "97fe4f65fbd6e9f7eb1eea80b144d1487f2a9b0aeae5fcf6f43b41491641884e")
    assert orc.pt_mul((GX, GY), 0x1E9AD) == pub                                        # 1_9_7File.pb:189
    pub = orc.parse_pubkey("036d05521c67b9cc1c0ef906b42215c7120c7302c34d9316a2726199bedac50936")
    assert orc.pt_mul((GX, GY), 0x16f7027bbf8454a5c) == pub                             # 1_9_7File.pb:191
    pub40 = orc.parse_pubkey("03a2efa402fd5268400c77c20e574ba86409ededee7c4020e4b9f0edbee53de0d4")
    assert orc.pt_mul((GX, GY), 0xE9AE4933D6) == pub40
    pub65 = orc.parse_pubkey("0230210c23b1a047bc9bdbb13448e67deddc108946de6de639bcc75d47c0216b1b")
    assert orc.pt_mul((GX, GY), 0x1a838b13505b26867) == pub65
    pub80 = orc.parse_pubkey("037e1238f7b1ce757df94faa9a2eb261bf0aeb9f84dbf81212104e78931c2a19dc")
    assert orc.pt_mul((GX, GY), 0xea1a5c66dcc11b5ad180) == pub80


def test_random_against_bigint(orc):
    rnd = random.Random(1234)
    edge = [0, 1, 2, P - 1, P - 2, 2**255, 2**256 - 2**32 - 978, 0x1000003D1, 2**128 - 1, 2**192]
    vals = edge + [rnd.randrange(P) for _ in range(300)]
    for a in vals:
        b = vals[rnd.randrange(len(vals))]
        assert orc.fe2("orc_fe_mul", a, b) == a * b % P
        assert orc.fe2("orc_fe_add", a, b) == (a + b) % P
        assert orc.fe2("orc_fe_sub", a, b) == (a - b) % P
        assert orc.fe1("orc_fe_sqr", a) == a * a % P
        assert orc.fe1("orc_fe_neg", a) == (-a) % P
        assert orc.fe1("orc_fe_inv", a) == (pow(a, -1, P) if a else 0)
    for _ in range(20):
        k = rnd.randrange(1, pyec.N)
        assert orc.pt_mul((GX, GY), k) == pyec.mul(pyec.G, k)


def test_y_from_x_and_parse_w(orc):
    assert orc.y_from_x(GX) in (GY, P - GY)
    assert orc.parse_w("22") == 4194304 and orc.parse_w("30") == 1073741824
    assert orc.parse_w("30.5") == 1518500249 and orc.parse_w("31.5") == 3037000499
    assert orc.parse_w("1000000") == 1000000

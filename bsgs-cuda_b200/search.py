"""Window enumeration and hit decoding around the GPU engine -- the host roles of the reference:
main-loop set-up (1_9_7File.pb:4689-4765, 5042-5064), GetJob (2077-2092), the per-GPU loop of
cuda() (2402-2523) and checkerThread (3933-4296), on Python ints.  Multi-GPU: rank r of R owns the
windows j = r, r+R, ... (static striding)."""
from __future__ import annotations
from . import ec
from . import formats


class Search:
    def __init__(self, T, B, p, w, htsz, pk, target_xy, start_key=1, rank=0, world=1):
        self.T, self.B, self.p, self.w, self.htsz = T, B, p, w, htsz
        self.N = T * B * p
        self.pk = pk
        self.target = target_xy
        self.rank, self.world = rank, world
        self.start_key = start_key
        self.addpubg = ec.neg(ec.mul(ec.G, 2 * w))                     # ADDPUBG  = -(2w)G
        self.centerbig = p * w                                         # CenterBig = p*w
        self.stride = 4 * self.N * w                                   # PRKADDBIG = 4*N*w
        self.pubaddbig = ec.neg(ec.mul(ec.G, self.stride))             # PUBADDBIG = -(4Nw)G
        self.findpub = ec.add(target_xy, ec.neg(ec.mul(ec.G, pk)))     # FINDPUBG = target - pk*G
        self.key0 = start_key + rank * self.stride
        # GlobPub = FINDPUBG - key*G - (p*w)G, done with two adds like the reference (5054-5063)
        self.P0 = ec.add(ec.add(self.findpub, ec.neg(ec.mul(ec.G, self.key0))), ec.neg(ec.mul(ec.G, self.centerbig)))
        self.D = ec.neg(ec.mul(ec.G, self.stride * world)) if world > 1 else self.pubaddbig
        self.rank_stride = self.stride * world

    def window_key(self, j):
        """key_j of this rank's j-th window."""
        return (self.key0 + j * self.rank_stride) % ec.N

    def window_point(self, j):
        return ec.add(self.P0, ec.mul(self.D, j)) if j else self.P0

    def windows_needed(self, pke):
        """The reference tests key_j > pke - pk AFTER the launch of window j (1_9_7File.pb:2512-2518), so it
        launches every window with key_j <= width and then one more: window j only covers
        key_j + p*w +- (2Nw + w), and the tail of the range lies in that last window."""
        width = pke - self.pk
        total = (width - self.start_key) // self.stride + 2 if width >= self.start_key else 1
        return (total - self.rank + self.world - 1) // self.world if total > self.rank else 0

    # ---- checkerThread
    def _try(self, K):
        K %= ec.N
        if K == 0:
            return None
        t = ec.mul(ec.G, K)
        if t is None or self.findpub is None or t[0] != self.findpub[0]:
            return None
        if t[1] != self.findpub[1]:
            K = ec.N - K
        fin = (self.pk + K) % ec.N
        t = ec.mul(ec.G, fin)
        if t is None or t[0] != self.target[0]:
            return None
        if t[1] != self.target[1]:
            fin = ec.N - fin
        return fin

    def check_verified(self, code, idx, j, payload):
        """A record that survived device-side verification: the baby-step index b = payload is known, so
        the candidates of checkerThread (1_9_7File.pb:4025-4124) are tried directly, without the htCPU table."""
        base = self.window_key(j) + self.centerbig
        giant = (idx + 1) * 2 * self.w
        if code == 4 or payload == 0xFFFFFFFE:
            return self._try(base + giant) or self._try(base - giant)
        b = payload + 1
        if code == 5:
            return self._try(base + b) or self._try(base - b)
        sgn = 1 if code == 1 else -1
        return self._try(base + sgn * (giant + b)) or self._try(base + sgn * (giant - b))

    def check_hit(self, htcpu, code, idx, j):
        """Turn one record of this rank's window j into the private key (or None: false positive)."""
        base = self.window_key(j) + self.centerbig
        giant = (idx + 1) * 2 * self.w
        if code == 4:                                                  # 1_9_7File.pb:3974-4019
            return self._try(base + giant) or self._try(base - giant)
        Pj = self.window_point(j)
        if code == 5:                                                  # 1_9_7File.pb:4025-4077
            for b in formats.htcpu_lookup(htcpu, self.htsz, Pj[0] & (2**64 - 1)):
                r = self._try(base + b + 1) or self._try(base - b - 1)
                if r:
                    return r
            return None
        if code not in (1, 2):
            return None
        Q = ec.mul(self.addpubg, idx + 1)                              # 1_9_7File.pb:4083-4124
        if code == 2:
            Q = ec.neg(Q)
        R = ec.add(Pj, Q)
        if R is None:
            return None
        sgn = 1 if code == 1 else -1
        for b in formats.htcpu_lookup(htcpu, self.htsz, R[0] & (2**64 - 1)):
            for s2 in (1, -1):
                r = self._try(base + sgn * (giant + s2 * (b + 1)))
                if r:
                    return r
        return None


def solve(engine, htcpu, T, B, p, w, htsz, pk, pke, target_xy, batch=16, rank=0, world=1, max_windows=None):
    """Search [pk, pke] for the key of target_xy with `engine` (tables already uploaded).
    Returns (key or None, windows_done, records_seen)."""
    s = Search(T, B, p, w, htsz, pk, target_xy, rank=rank, world=world)
    nwin = s.windows_needed(pke)
    if max_windows is not None:
        nwin = min(nwin, max_windows)
    done, seen = 0, 0
    while done < nwin:
        chunk = min(nwin - done, 64 * batch)
        hits, d = engine.run(s.window_point(done), s.D, chunk, batch=min(batch, chunk), stop_after_hits=1)
        for (code, idx, win), payload in zip(hits, list(engine.payloads)):
            seen += 1
            if getattr(engine, "verify", False) and payload != 0xFFFFFFFF:
                key = s.check_verified(code, idx, done + win, payload)
            else:
                key = s.check_hit(htcpu, code, idx, done + win)
            if key is not None:
                return key, done + d, seen
        done += d
    return None, done, seen

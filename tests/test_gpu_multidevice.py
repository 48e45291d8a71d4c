"""Multi-GPU through the product's OWN host model -- one process, one context per device, like the
reference's CreateThread(@cuda(), i) workers each holding a replica of the tables
(1_9_7File.pb:2337-2350, 4776).  Skipped on a one-GPU box; run with `gpurun --gpus 2`."""
import os
import random
import subprocess
import pytest
from tests import pyec
from tests.helpers import norm_hits, random_point

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "bsgs-cuda_b200", "bsgs_b200")


@pytest.fixture(scope="module")
def pkg():
    import bsgs_cuda_b200 as b
    b.native.load()
    return b


def _ndev(pkg):
    return len(pkg.native.devices())


def test_bcast_ht_across_real_devices(pkg):
    """bsgs_bcast_ht from device 0 to every other device: images (htGPU and htCPU form) equal, the same
    records from every device, verified records included; contexts driven from one thread each."""
    n = _ndev(pkg)
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    import threading
    T, B, p, w, htsz = 64, 8, 16, 1 << 20, 18
    rnd = random.Random(2)
    es = [pkg.Engine(T, B, p, w, htsz, device=d) for d in range(n)]
    try:
        for e in es:
            e.giants_generate()
        es[0].ht_generate()
        pkg.native.bcast_ht(es)
        g0, c0 = bytes(es[0].ht_export("gpu")), bytes(es[0].ht_export("cpu"))
        for e in es[1:]:
            assert bytes(e.ht_export("gpu")) == g0 and bytes(e.ht_export("cpu")) == c0
        N = T * B * p
        pts = [pyec.mul(pyec.G, (k + 1) * 2 * w + b) for k, b in ((0, 1), (N - 1, w), (N // 2, 12345))] + [random_point(rnd)]
        for e in es:
            e.set_verify(True)
        out = [None] * n

        def work(i):
            out[i] = [norm_hits(es[i].step(pt)) for pt in pts]
        th = [threading.Thread(target=work, args=(i,)) for i in range(n)]
        [t.start() for t in th]; [t.join() for t in th]
        assert (1, 0) in out[0][0] and (1, N - 1) in out[0][1] and (1, N // 2) in out[0][2]
        assert all(o == out[0] for o in out)
    finally:
        for e in es:
            e.close()


def test_cpp_host_on_two_devices(pkg, tmp_path):
    """The C++ host with -d 0,1: tables built once on device 0 and copied with bsgs_bcast_ht, the windows
    striped over the two workers (j = r, r+2, ...), keys of several targets recovered whichever worker
    owns their window."""
    if _ndev(pkg) < 2:
        pytest.skip("needs at least 2 GPUs")
    if not os.path.exists(EXE):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "bsgs-cuda_b200"), "bsgs_b200"])
    rnd = random.Random(9)
    T, B, p, wl, htsz = 32, 4, 16, 14, 12
    stride = 4 * T * B * p * (1 << wl)
    pk = 0x2000000000000
    secrets = [pk + j * stride + rnd.randrange(stride) for j in (0, 1, 2, 5)]        # windows of both workers
    with open(tmp_path / "pubs.txt", "w") as f:
        for s in secrets:
            f.write(pkg.ec.compress(pyec.mul(pyec.G, s)) + "\n")
    args = f"-t {T} -b {B} -p {p} -w {wl} -htsz {htsz} -pk {pk:x} -pke {pk + 8 * stride:x} -infile pubs.txt -nosave -d 0,1".split()
    r = subprocess.run([EXE] + args, capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert r.returncode == 0 and "found 4 of 4" in r.stdout, r.stdout
    for i, s in enumerate(secrets, start=1):
        assert f"KEY[{i}]: 0x{s:064x}" in r.stdout

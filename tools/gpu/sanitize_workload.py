"""Small workload for compute-sanitizer (tools/gpu/sanitize.sh): every kernel family of the library at orders where the tile
loops, the table double-buffering and the TMA / mbarrier hand-offs are all exercised, checked against the oracle."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_lib as ol          # noqa: E402
import fastecc_b200 as fe        # noqa: E402

fe.init(0)
o = ol.load_oracle()
for L, S in ((11, 256), (12, 128), (13, 64), (10, 16), (6, 1024)):
    N = 1 << L
    a = ol.fill_B(o, N, S)
    t = torch.from_numpy(a.view(np.int32)).cuda()
    fe.rs_encode_dev(t)
    assert np.array_equal(t.cpu().numpy().view(np.uint32), ol.o_encode(o, a)), ("encode", L, S)
    t = torch.from_numpy(a.view(np.int32)).cuda()
    fe.ntt_dev(t, False)
    assert np.array_equal(t.cpu().numpy().view(np.uint32), ol.o_ntt(o, a, False)), ("ntt", L, S)
    if L >= 11:
        t = torch.from_numpy(a.view(np.int32)).cuda()
        fe.rs_encode_asym_dev(t, N // 2)
        assert np.array_equal(t[:N // 2].cpu().numpy().view(np.uint32), ol.o_encode(o, a)[::2]), ("asym", L, S)
    b = a.copy(); fe.EncodeReedSolomon_body(b, N, S)
    assert np.array_equal(b, ol.o_encode(o, a)), ("host", L, S)
raw = torch.randint(0, 256, (64, 4096), dtype=torch.uint8, device="cuda")
w = fe.bytes_to_gfp_dev(raw)
assert bool((fe.gfp_to_bytes_dev(w, 1024) == raw).all())
torch.cuda.synchronize()
print("sanitize workload ok: %d kernels launched" % fe.kernel_launches())

"""Row-sharded fold step on ONE GPU: the ranks are played one after another, their partial buffers are
concatenated the way an all_gather would leave them, and the combined commitments, folded witness and folded error
vector must equal the unsharded step's, bit for bit (which tests/test_gpu_witness.py and smoke() pin to the oracle)."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("world", [2, 3, 8])
def test_sharded_fold_step_equals_unsharded(world):
    import fold_step as F
    from mira_b200.sharding import shard_range
    k = 10
    whole = F.GpuFoldStep(k)
    want = whole.step(False)
    torch.cuda.synchronize()
    bufs, ranks = [], []
    for rank in range(world):
        r = F.ShardedGpuFoldStep(k, rank, world)
        buf = r.step()                      # queued on the rank's own stream
        r.stream.synchronize()
        bufs.append(buf.clone())
        ranks.append(r)
    torch.cuda.synchronize()
    gathered = torch.cat(bufs)
    got = ranks[0].combine(gathered, world)
    assert got == want
    assert len(got) == ranks[0].n_commits == 13
    # the folded vectors are the same rows of the unsharded result
    for ci, s in enumerate(whole.sh):
        rows, cols = s["rows"], s["meta"]["num_advice"]
        w_whole = whole.state[ci]["W_out"].view(cols, rows, 32)
        e_whole = whole.state[ci]["E_out"].view(rows, 32)
        for rank, r in enumerate(ranks):
            lo, hi = shard_range(rows, world, rank)
            assert torch.equal(r.state[ci]["W_out"].view(cols, hi - lo, 32), w_whole[:, lo:hi])
            assert torch.equal(r.state[ci]["E_out"].view(hi - lo, 32), e_whole[lo:hi])


def test_partial_batch_device_matches_host_partials():
    """mira_msm_partial_batch_dev / mira_msm_combine_dev against the host-memory partial / combine pair."""
    import gpu_util
    import pyref as R
    from mira_b200 import CommitmentKey, combine_partials, combine_partials_device
    curve, n = R.GRUMPKIN, 5000
    halves = []
    for part in range(2):
        bases = gpu_util.gen_bases_dev(curve, 77, n, first=part * n)
        ck = CommitmentKey(curve, bases, on_device=True)
        vecs = [gpu_util.gen_scalars_dev(curve, 78 + v, n, v % 2, first=part * n) for v in range(3)]
        out = torch.zeros(3 * 128, dtype=torch.uint8, device="cuda")
        ck.partial_batch_device([v.data_ptr() for v in vecs], n, out.data_ptr())
        one = torch.zeros(128, dtype=torch.uint8, device="cuda")
        ck.partial_batch_device([vecs[1].data_ptr()], n, one.data_ptr())
        torch.cuda.synchronize()
        host = [ck.partial(v.data_ptr(), n, on_device=True) for v in vecs]
        halves.append((out, one, host))
    gathered = torch.cat([halves[0][0], halves[1][0]])
    got = combine_partials_device(curve, gathered.data_ptr(), 2, 3, 3 * 128)
    want = [combine_partials(curve, halves[0][2][v] + halves[1][2][v]) for v in range(3)]
    assert got == want
    single = combine_partials_device(curve, torch.cat([halves[0][1], halves[1][1]]).data_ptr(), 2, 1, 128)
    assert single == [want[1]]

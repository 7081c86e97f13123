"""CPU, world_size 2 over gloo: the N>1 path of the benchmark / pipeline -- frames sharded across
ranks with no data-path collective, timing reduced with max-over-ranks, per-square records gathered
to rank 0.  Compute is the oracle-backed fake engine (no GPU here)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions():
    from chessboard_vision_b200.sharding import shard_range, stream_owner
    for n in (0, 1, 7, 64, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    assert [stream_owner(s, 8) for s in range(10)] == [0, 1, 2, 3, 4, 5, 6, 7, 0, 1]


def test_rank_placement_and_proportional_split():
    from chessboard_vision_b200.sharding import place_rank, proportional_split
    assert [place_rank(r, 8, 8) for r in range(8)] == list(range(8))           # whole node: identity
    assert [place_rank(r, 4, 8) for r in range(4)] == [4, 5, 6, 7]             # partial node: the GPUs with the fast host path
    assert [place_rank(r, 2, 8) for r in range(2)] == [6, 7] and place_rank(0, 1, 8) == 7
    assert [place_rank(r, 4, 4) for r in range(4)] == [0, 1, 2, 3]             # launcher shows only the job's GPUs
    assert place_rank(0, 1, 1) == 0
    shares = proportional_split(2048, [23.4] * 4 + [35.5] * 4, multiple=8)
    assert sum(shares) == 2048 and all(x % 8 == 0 for x in shares)
    assert shares[0] < 256 < shares[7] and abs(shares[7] / shares[0] - 35.5 / 23.4) < 0.08
    assert proportional_split(10, [1, 1, 1], multiple=4) in ([4, 4, 2], [4, 2, 4], [2, 4, 4]) or sum(proportional_split(10, [1, 1, 1], 4)) == 10
    assert proportional_split(7, [5.0], 8) == [7]


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from chessboard_vision_b200 import synth
    from chessboard_vision_b200.engine import grid_rects, SQ_PD_STATS, SQ_CD_CALIBRATE, SQ_CD_DETECT
    from chessboard_vision_b200.sharding import shard_range, dist_max, gather_records
    from fake_engine import FakeEngine
    import oracle as O
    eng = FakeEngine()
    n_total, S = 5, 80
    lo, hi = shard_range(n_total, rank, world)
    rects, _ = grid_rects(S)
    state = eng.new_state(hi - lo, S, S)
    boards = np.stack([O.warp(synth.board_frame(90, 120, s), O.get_perspective(synth.calib_points(90, 120), [[0, 0], [S, 0], [0, S], [S, S]]), S)
                       for s in range(lo, hi)])
    p = eng.square_params(ops=SQ_PD_STATS | SQ_CD_CALIBRATE | SQ_CD_DETECT)
    local = eng.squares(boards, rects, p, state)
    dist.barrier()
    slowest = dist_max(10.0 + rank)                # every rank learns the slowest rank's time
    assert slowest == 10.0 + world - 1
    allrec = gather_records(local, dst=0)
    if rank == 0:
        np.save(os.path.join(out_dir, "gathered.npy"), allrec)
    else:
        assert allrec is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_one(tmp_path):
    torch = pytest.importorskip("torch")
    import torch.multiprocessing as mp
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "gathered.npy"))
    # single-process run of the same 5 frames
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from chessboard_vision_b200 import synth
    from chessboard_vision_b200.engine import grid_rects, SQ_PD_STATS, SQ_CD_CALIBRATE, SQ_CD_DETECT
    from fake_engine import FakeEngine
    import oracle as O
    eng = FakeEngine()
    S = 80
    rects, _ = grid_rects(S)
    M = O.get_perspective(synth.calib_points(90, 120), [[0, 0], [S, 0], [0, S], [S, S]])
    boards = np.stack([O.warp(synth.board_frame(90, 120, s), M, S) for s in range(5)])
    ref = eng.squares(boards, rects, eng.square_params(ops=SQ_PD_STATS | SQ_CD_CALIBRATE | SQ_CD_DETECT), eng.new_state(5, S, S))
    assert got.shape == ref.shape == (5, 64)
    assert got.tobytes() == ref.tobytes()

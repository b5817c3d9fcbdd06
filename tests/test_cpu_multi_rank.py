"""N > 1 path of bench.py on CPU: world_size 2, gloo, 127.0.0.1.  The hot path shards by independent stream (no
data-path collective), so what has to be right across ranks is: distinct streams per rank, barrier + max-over-ranks
timing, whole-job aggregation, and that the reference arm lets rank 0 alone do the work."""
import os
import socket
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from libde265_b200 import shard, synth
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # every rank decodes its own stream: different seeds -> different records
    pic = synth.make_picture(64, 64, "P", seed=shard.stream_seed(rank), dst_slot=1, ref_slots=(0,))
    sig = int(pic.pus["mv"].astype("int64").sum()) * 31 + len(pic.tus)
    sigs = [None] * world
    dist.all_gather_object(sigs, sig)
    dist.barrier()
    fps, ms = shard.job_frames_per_second(32 * 4, 100.0 * (rank + 1))  # rank 1 is twice as slow
    out.put((rank, sigs, fps, ms))
    dist.destroy_process_group()


def test_two_rank_sharding_and_aggregation():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for rank, sigs, fps, ms in res:
        assert sigs[0] != sigs[1], "ranks must decode different streams"
        assert ms == 200.0, "device time = max over ranks"
        assert abs(fps - (32 * 4 * 2) / 0.2) < 1e-6, "whole-job frames/s over all ranks"


def test_reference_arm_only_rank0_works():
    """bench.py --impl reference under a 2-rank launch: rank 1 exits 0 without output, rank 0 prints the JSON line."""
    env = dict(os.environ, WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    r1 = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                        env=dict(env, RANK="1", LOCAL_RANK="1"), capture_output=True, text=True, timeout=120)
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_reference_arm_json_line():
    """The CPU arm's JSON line: same metric / unit / config keys as the GPU arm, impl = reference, a cpu_baseline that says
    what was timed, e2e with zero transfer bytes (small picture size so the test is quick)."""
    import json
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--width", "256",
                        "--height", "128"], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    assert line["impl"] == "reference" and line["metric"] == base["metric"] and line["unit"] == "frames/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("port", "reference") and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb and cb["frames_per_s_per_core"] > 0 and "one_core_alone" in cb
    assert "workload" in line["config"] and "model" not in line["config"]

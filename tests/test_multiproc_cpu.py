"""CPU, world_size 2 over gloo: the host-side N>1 logic (sharding, routing, weight broadcast wrapper, max-over-ranks
aggregation) that bench.py runs over NCCL on the GPU box."""
import os
import socket
import subprocess
import sys
import textwrap

from helix_b200 import replica

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 32, 100001):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                a, b = replica.shard_range(n, world, r)
                seen += list(range(a, b)) if n < 1000 else [a, b]
            if n < 1000:
                assert seen == list(range(n))
            else:
                assert seen[0] == 0 and seen[-1] == n and all(seen[2 * i + 1] == seen[2 * i + 2] for i in range(world - 1))


def test_route_least_active_matches_scheduler_rule():
    assert replica.route_least_active([0, 0, 0, 0], 6) == [0, 1, 2, 3, 0, 1]
    assert replica.route_least_active([3, 1, 2], 4) == [1, 1, 2, 0]
    # 256 sessions over 8 idle replicas -> 32 each (config 5 of BASELINE.json)
    r = replica.route_least_active([0] * 8, 256)
    assert [r.count(i) for i in range(8)] == [32] * 8


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, %r)
    import torch, torch.distributed as dist
    from helix_b200 import replica
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    # weight arena stand-in: replicas start different, one broadcast makes them identical to rank 0
    g = torch.Generator().manual_seed(1234 + rank)
    arena = torch.randint(0, 256, (1 << 20,), dtype=torch.uint8, generator=g)
    want = torch.randint(0, 256, (1 << 20,), dtype=torch.uint8, generator=torch.Generator().manual_seed(1234))
    assert (rank == 0) == bool(torch.equal(arena, want))
    replica.broadcast_buffer(dist, arena, src=0)
    assert torch.equal(arena, want)
    # units sharded without overlap; aggregate = sum of units / max of times
    a, b = replica.shard_range(101, world, rank)
    units, secs, rate = replica.aggregate_throughput(dist, float(b - a), 1.0 + rank)
    assert units == 101.0 and secs == float(world) and abs(rate - 101.0 / world) < 1e-12
    routes = replica.route_least_active([0] * world, 10)
    gathered = [None] * world
    dist.all_gather_object(gathered, routes)
    assert all(x == gathered[0] for x in gathered)   # every rank derives the same routing
    dist.destroy_process_group()
    sys.stdout.write("rank" + str(rank) + "-ok" + chr(10))
    sys.stdout.flush()
""")


def test_gloo_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "rank0-ok" in out.stdout and "rank1-ok" in out.stdout, out.stdout

"""Native auto-growth best-fit allocator (csrc/runtime/allocator.cpp) driven through its host backend: best fit, split, coalesce,
chunk growth and release, peaks, out-of-memory recovery.  Parity model: test/cpp/fluid/memory (auto_growth_best_fit_allocator_test)."""
import pytest

from paddle_b200._build import load

m = load()
pytestmark = pytest.mark.skipif(m is None or not hasattr(m, "AutoGrowthAllocator"), reason="native extension not built")

MB = 1 << 20


def test_alignment_split_and_coalesce():
    a = m.AutoGrowthAllocator("host", MB, 256)
    p1, p2, p3 = a.alloc(1000), a.alloc(5000), a.alloc(300)
    assert p1 % 256 == 0 and p2 % 256 == 0 and p3 % 256 == 0
    assert p2 == p1 + 1024 and p3 == p2 + 5120                      # carved front to back from one chunk
    s = a.stats()
    assert s["num_chunks"] == 1 and s["reserved"] == MB and s["allocated"] == 1024 + 5120 + 512 and s["num_backend_allocs"] == 1
    a.free(p2)
    assert a.alloc(4000) == p2                                      # the hole is reused (best fit), its tail split off
    a.free(p1)
    a.free(p2)
    a.free(p3)
    s = a.stats()
    assert s["allocated"] == 0 and a.largest_free_block() == MB      # everything merged back into one block
    assert s["num_merges"] >= 3 and s["allocated_peak"] == 1024 + 5120 + 512


def test_best_fit_prefers_smallest_block_that_fits():
    a = m.AutoGrowthAllocator("host", MB, 256)
    ps = [a.alloc(s) for s in (64 * 1024, 1024, 16 * 1024, 1024, 256 * 1024, 1024)]
    a.free(ps[0])      # 64 KB hole
    a.free(ps[2])      # 16 KB hole
    a.free(ps[4])      # 256 KB hole
    assert a.alloc(10 * 1024) == ps[2]                               # 16 KB hole, not the first (64 KB) or the largest
    assert a.alloc(60 * 1024) == ps[0]
    assert a.alloc(200 * 1024) == ps[4]


def test_growth_large_requests_and_release_idle():
    a = m.AutoGrowthAllocator("host", MB, 256)
    small = a.alloc(1024)
    big = a.alloc(5 * MB)                                            # larger than a chunk: gets its own chunk of exactly that size
    s = a.stats()
    assert s["num_chunks"] == 2 and s["reserved"] == 6 * MB
    a.free(big)
    assert a.release_idle() == 5 * MB                                # the idle chunk returns to the backend, the busy one stays
    s = a.stats()
    assert s["num_chunks"] == 1 and s["reserved"] == MB and s["reserved_peak"] == 6 * MB and s["num_backend_frees"] == 1
    a.free(small)
    assert a.release_idle() == MB and a.stats()["reserved"] == 0
    a.reset_peak()
    assert a.stats()["reserved_peak"] == 0


def test_out_of_memory_releases_idle_chunks_then_raises():
    a = m.AutoGrowthAllocator("host", MB, 256)
    a.set_host_limit(3 * MB)
    p = [a.alloc(MB) for _ in range(3)]
    with pytest.raises(MemoryError):
        a.alloc(MB)
    a.free(p[0])
    a.free(p[1])                                                      # two idle 1 MB chunks: not adjacent, a 2 MB request fits neither
    q = a.alloc(2 * MB)                                               # ... so they are handed back and one 2 MB chunk is made
    s = a.stats()
    assert s["reserved"] == 3 * MB and s["num_backend_frees"] == 2 and q
    with pytest.raises(RuntimeError):
        a.free(12345)                                                 # not ours


def test_many_random_allocations_stay_consistent():
    import random

    rng = random.Random(0)
    a = m.AutoGrowthAllocator("host", 4 * MB, 512)
    live = {}
    for step in range(4000):
        if live and (rng.random() < 0.45 or len(live) > 200):
            ptr = rng.choice(list(live))
            a.free(ptr)
            del live[ptr]
        else:
            n = rng.choice([1, 100, 4096, 70000, 300000, 2 * MB])
            ptr = a.alloc(n)
            assert ptr % 512 == 0
            size = (max(n, 1) + 511) // 512 * 512
            for q, sz in live.items():                                # no overlap with any live block
                assert ptr + size <= q or q + sz <= ptr
            live[ptr] = size
    s = a.stats()
    assert s["allocated"] >= sum(live.values())                       # blocks may keep an unsplittable tail
    for ptr in list(live):
        a.free(ptr)
    assert a.stats()["allocated"] == 0
    a.release_idle()
    assert a.stats()["reserved"] == 0

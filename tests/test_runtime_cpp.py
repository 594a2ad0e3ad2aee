"""C++ sampler runtime (csrc/runtime.cpp): block manager + scheduler, exercised on CPU."""
import pytest

_C = pytest.importorskip("nanorlhf_b200._C")


def test_block_manager_refcounts():
    bm = _C.BlockManager(8, 16)
    a = bm.allocate(3)
    assert bm.num_free() == 5 and bm.blocks_for(17) == 2
    bm.incref(a[:2])
    bm.release(a)
    assert bm.num_free() == 6                      # two pages still referenced
    bm.release(a[:2])
    assert bm.num_free() == 8
    with pytest.raises(RuntimeError):
        bm.release(a[:1])
    with pytest.raises(RuntimeError):
        bm.allocate(9)


def test_prefix_sharing_and_reserve_policy():
    s = _C.Scheduler(64, 16, 16, True)
    g = s.add_request(40, 4, 30)                   # 40-token prompt: 2 full shared pages + private tail
    assert s.admit() == [g]
    seqs = s.group_seqs(g)
    tables = [s.block_table(i) for i in seqs]
    assert all(t[:2] == tables[0][:2] for t in tables) and len({t[2] for t in tables}) == 4
    assert all(len(t) == 5 for t in tables)        # ceil((40+30)/16) = 5 pages reserved per sample
    bt = s.block_tables(seqs, 7, -1)               # the whole batch's page table in one call (what the sampler uploads)
    assert bt.shape == (4, 7) and bt.dtype.name == "int32"
    assert [row[:5].tolist() for row in bt] == tables and (bt[:, 5:] == -1).all()
    import pytest
    with pytest.raises(RuntimeError):
        s.block_tables(seqs, 3, -1)                # narrower than a sequence's page list
    assert s.num_free_blocks() == 64 - (2 + 4 * 3)
    tok, slot = s.prefill_slots(g)
    assert len(tok) == 32 + 4 * 8 and tok[:32] == list(range(32))
    s.finish(seqs)
    assert s.num_free_blocks() == 64 and s.num_running_seqs() == 0


def test_admission_waits_for_pages_then_proceeds():
    s = _C.Scheduler(10, 16, 64, True)
    g0 = s.add_request(16, 1, 100)                 # 8 pages
    g1 = s.add_request(16, 1, 100)
    assert s.admit() == [g0] and s.num_waiting() == 1
    assert s.admit() == []
    s.finish(s.group_seqs(g0))
    assert s.admit() == [g1]


def test_on_demand_preemption_and_recompute():
    s = _C.Scheduler(6, 16, 64, False)
    g0 = s.add_request(16, 1, 200)
    g1 = s.add_request(16, 1, 200)
    assert s.admit() == [g0, g1]                   # 2 pages each at admission
    pre = []
    for _ in range(80):
        live = [q for g in s.running_groups() for q in s.group_seqs(g)]
        pre += s.advance(live)
        if pre:
            break
    assert pre == [g1]                             # the youngest group is the victim
    assert s.num_waiting() == 1 and s.running_groups() == [g0]
    s.finish(s.group_seqs(g0))
    assert s.admit() == [g1] and s.admissions(g1) == 2


def test_runtime_under_address_and_ub_sanitizers():
    """the host runtime built with -fsanitize=address,undefined survives a randomized
    admit / advance / finish / preempt workload with page accounting checked every step."""
    import os
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        import pytest
        pytest.skip("no host compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run(["bash", os.path.join(root, "bench", "asan_runtime.sh")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all 257 pages returned" in r.stdout

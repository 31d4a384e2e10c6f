"""The N>1 host logic on CPU: world_size-2 (and 3) `gloo` process groups.  Each rank scans its own shard --
with the oracle standing in for the device scan, there is no GPU here -- and the gathered, rebased lists must
equal the oracle's answer on the whole text, for newline and paragraph ($$) records."""
import os, socket
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import _oracle, _corpus
from agrep_b200 import shard


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, text, pattern, kw, delim_bytes, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cuts = shard.cut_points(text, world, delim_bytes)
        part = text[cuts[rank]:cuts[rank + 1]]
        a = _oracle.compile(pattern, **kw)
        cnt, recs = _oracle.scan(a, part)
        t = torch.zeros((max(cnt, 1), 4), dtype=torch.int64)
        for i, (b, e, j) in enumerate(recs):
            t[i, 0], t[i, 1], t[i, 2] = b, e, j
        allr = shard.gather_records(t, cnt, cuts[rank], dist, closes=shard.count_closes(part, delim_bytes), delim=delim_bytes, shard_head=part[:8])
        total = torch.tensor([cnt]); dist.all_reduce(total)
        if rank == 0:
            q.put((int(total), [(int(x[0]), int(x[1]), int(x[2])) for x in allr]))
    finally:
        dist.destroy_process_group()


def run_world(world, text, pattern, kw, delim_bytes):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, text, pattern, kw, delim_bytes, q)) for r in range(world)]
    [p.start() for p in ps]
    out = q.get(timeout=120)
    [p.join(timeout=60) for p in ps]
    assert all(p.exitcode == 0 for p in ps)
    return out


@pytest.mark.parametrize("world", [2, 3])
def test_newline_shards_gather_to_the_whole_answer(world):
    text = _corpus.make_text(3000, seed=21)
    kw = dict(k=2, linenum=1)
    a = _oracle.compile("because each", **kw)
    cnt, recs = _oracle.scan(a, text)
    total, got = run_world(world, text, "because each", kw, b"\n")
    assert total == cnt
    # a shard's first record begins at its virtual '\n' (-1 + base) = the real '\n' before it in the whole text
    assert [(b, e) for b, e, _ in got] == [(b, e) for b, e, _ in recs]
    # -n ordinals: shard-local j + the closes of the shards before (an exclusive prefix sum over the ranks)
    assert [j for _, _, j in got] == [j for _, _, j in recs]


def test_paragraph_shards():
    text = _corpus.make_text(2500, seed=22, paragraphs=True)
    kw = dict(k=1, linenum=1, delim="$$")
    a = _oracle.compile("state", **kw)
    cnt, recs = _oracle.scan(a, text)
    total, got = run_world(2, text, "state", kw, b"\n\n")
    assert total == cnt
    assert [e for _, e, _ in got] == [e for _, e, _ in recs]   # record ends are global; begins differ by the leading delimiter
    assert [j for _, _, j in got] == [j for _, _, j in recs]


def test_user_delimiter_seams_keep_the_opening_delimiter():
    """with -d the record after a cut begins at the delimiter that closed the record before it, as in the whole text"""
    text = _corpus.make_text(1200, seed=24).replace(b"\n", b"; ")
    kw = dict(k=1, linenum=1, delim="; ")
    a = _oracle.compile("people", **kw)
    cnt, recs = _oracle.scan(a, text)
    total, got = run_world(2, text, "people", kw, b"; ")
    assert total == cnt and cnt > 5
    assert [(b, e) for b, e, _ in got] == [(b, e) for b, e, _ in recs]


def test_cut_points_are_record_aligned():
    text = _corpus.make_text(500, seed=23, paragraphs=True)
    for world in (2, 4, 7):
        cuts = shard.cut_points(text, world, b"\n")
        assert cuts[0] == 0 and cuts[-1] == len(text) and cuts == sorted(cuts)
        assert all(text[c - 1] == 10 for c in cuts[1:-1])
    assert shard.page_shards(64 * 4096, 8) == [(r * 8 * 4096, 8 * 4096) for r in range(8)]


@pytest.mark.parametrize("delim", ["aba", "e e", "$$x"])
def test_delimiter_rule_of_the_host_helper(delim):
    """shard._delim_ends_at (the host-side statement of the device's delimiter rule, used to cut texts in these tests) against
    the oracle's record ends, for a delimiter that overlaps itself: occurrences are taken from the left, overlapping ones dropped"""
    if delim == "$$x":
        delim, dbytes, data = "$$", b"\n\n", _corpus.make_text(300, seed=5, paragraphs=True) + b"\n" * 5 + b"tail state\n"
    else:
        dbytes, data = delim.encode(), _corpus.overlap_text(delim, 11)
    a = _oracle.compile("zzzzqqqq", k=0, linenum=1, inverse=1, delim=delim)          # every record
    cnt, recs = _oracle.scan(a, data)
    L = len(dbytes)
    ends = {e + L - 1 for _, e, _ in recs if e + L - 1 < len(data)}                # last byte of each closing delimiter inside the text
    mine = {q for q in range(len(data)) if shard._delim_ends_at(data, q, dbytes)}
    assert cnt > 50 and ends <= mine
    # closes the oracle does not list are records it drops (empty ones between adjacent delimiters), never a different parse
    for q in mine - ends:
        assert shard._delim_ends_at(data, q - L, dbytes) or q - L < 0

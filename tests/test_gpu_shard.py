"""The cut rule of the sharded scan on ONE GPU: the shards of a text are scanned one after the other through the local
half of agb_scan_sharded (agb_scan_shard_local: halos on either side, a record belongs to the shard that holds the last
byte of the delimiter that opened it) and stitched the way the gather does; the result must be the oracle's answer on the
whole text -- offsets, ordinals, counts, delimiter totals.  The cuts fall on multiples of 512 bytes, i.e. in the middle of
records; tests/shard_nccl_worker.py is the same over NCCL with one process per GPU."""
import ctypes as C
import random
import pytest
import _oracle, _corpus
import agrep_b200 as ag
from agrep_b200 import _lib

pytestmark = pytest.mark.gpu


def scan_in_shards(pattern, kw, data, world, want_ordinals=True, halo_right=_lib.HALO_RIGHT):
    import torch
    L = _lib.lib()
    p = ag.Pattern(pattern, **kw)
    n = len(data)
    per = max(512, (n // world) // 512 * 512)
    offs = [min(r * per, n) for r in range(world)] + [n]
    cap = n + 2
    out, closes_before, origin, n_closes, matched = [], 0, 0, 0, 0
    for r in range(world):
        n_local = offs[r + 1] - offs[r]
        hl = _lib.HALO_LEFT if r > 0 else 0
        hr = min(halo_right, n - offs[r + 1])
        ext = data[offs[r] - hl:offs[r + 1] + hr]
        t = torch.frombuffer(bytearray(ext + b"\0" * 64), dtype=torch.uint8).cuda()
        rec = torch.zeros((cap, 4), dtype=torch.int64, device="cuda")
        res, part = _lib.Result(), _lib.ShardPart()
        want = _lib.WANT_RECORDS | (_lib.WANT_ORDINALS if want_ordinals else 0)
        rc = L.agb_scan_shard_local(p._h, C.c_void_p(t.data_ptr() + hl), n_local, hl, hr, int(r == 0), int(r == world - 1),
                                    int(offs[r + 1] + hr >= n), want, C.c_void_p(rec.data_ptr()), cap, None, C.byref(res), C.byref(part))
        assert rc == 0, L.agb_last_error()
        if r == 0:
            origin = part.ord_origin
            n_closes += part.virt
        rows = rec[:res.n_records].cpu().tolist()
        base = offs[r] + part.byte_base
        for b, e, j, _ in rows:
            out.append((b + base, e + base, j + origin + closes_before - part.ord_fix))
        closes_before += part.closes
        n_closes += part.closes
        matched += res.n_matched
    return matched, out, n_closes


def ragged_text(seed, nlines=2500, sep="\n"):
    rnd = random.Random(seed)
    words = [w for w in _corpus.make_text(300, seed=1).decode().split() if w.isalpha()]
    lines = []
    for i in range(nlines):
        ln = rnd.choice([0, 0, 1, 5, 30, 60, 90, 255, 256, 257, 511, 512, 513, 700, 1500])
        row = ""
        while len(row) < ln:
            row += rnd.choice(words) + " "
        lines.append(row[:ln])
    return sep.join(lines).encode()


@pytest.mark.parametrize("world", [2, 3, 5])
@pytest.mark.parametrize("pattern,kw", [
    ("because each", dict(k=2, linenum=1)), ("the", dict()), ("people", dict(k=1, linenum=1, inverse=1)),
    ("t[hx]e", dict(k=0, linenum=1)), ("governmental", dict(k=3, linenum=1, nocase=1)), ("state;world", dict(k=1, linenum=1)),
])
def test_shards_of_newline_records(world, pattern, kw):
    data = ragged_text(5) + b"\n" + _corpus.make_text(1500, seed=9)
    a = _oracle.compile(pattern, **kw)
    cnt, recs = _oracle.scan(a, data)
    matched, got, n_closes = scan_in_shards(pattern, kw, data, world)
    assert matched == cnt and cnt > 0
    keep = (lambda t: t) if a.engine != 4 else (lambda t: t[:2])          # sgrep/bm has no j
    assert [keep(t) for t in got] == [keep(t) for t in recs]
    from agrep_b200 import shard
    assert n_closes == shard.count_closes(data, b"\n")


@pytest.mark.parametrize("world", [2, 4])
@pytest.mark.parametrize("delim,dbytes,kw", [("$$", b"\n\n", dict(k=1, linenum=1, wordbound=1)), ("; ", b"; ", dict(k=1, linenum=1)),
                                             ("$$", b"\n\n", dict(k=0, linenum=1, inverse=1)), ("aba", b"aba", dict(k=1, linenum=1))])
def test_shards_of_user_delimiters(world, delim, dbytes, kw):
    """paragraph records (runs of newlines that straddle the cuts: the left halo resolves the greedy pairing) and a
    2-byte delimiter whose bytes fall on either side of a cut"""
    if dbytes == b"\n\n":
        data = _corpus.make_text(2500, seed=31, paragraphs=True) + b"\n" * 7 + _corpus.make_text(800, seed=32, paragraphs=True)
    elif dbytes == b"aba":
        from _corpus import overlap_text      # chains of overlapping occurrences ("abababa") across the cuts
        data = overlap_text("aba", 7) + overlap_text("aba", 8)
    else:
        data = ragged_text(6, sep="; ")
    for shift in (0, 1, 3):                                  # move the text under the fixed cuts
        d = b"x" * shift + data
        a = _oracle.compile("state", delim=delim, **kw)
        cnt, recs = _oracle.scan(a, d)
        matched, got, n_closes = scan_in_shards("state", dict(delim=delim, **kw), d, world)
        assert matched == cnt and cnt > 0
        assert got == list(recs)


def test_a_record_longer_than_the_halo_is_an_error_not_a_wrong_answer():
    data = _corpus.make_text(40, seed=3) + b"y" * 3000 + b" because each " + b"z" * 3000 + b"\n" + _corpus.make_text(40, seed=4)
    with pytest.raises(AssertionError, match="halo"):
        scan_in_shards("because each", dict(k=1, linenum=1), data, 4, halo_right=512)
    matched, got, _ = scan_in_shards("because each", dict(k=1, linenum=1), data, 4)      # with the real halo it is fine
    a = _oracle.compile("because each", k=1, linenum=1)
    cnt, recs = _oracle.scan(a, data)
    assert matched == cnt and got == list(recs)


def test_sharded_api_with_one_rank_is_the_plain_scan():
    """agb_scan_sharded / agb_bestmatch_sharded over a world of one (NCCL communicator of size 1): same answers as the
    unsharded calls"""
    import torch
    from agrep_b200 import shard
    data = _corpus.make_text(3000, seed=12)
    comm = shard.Comm()
    buf, ptr = shard.shard_buffer(torch, len(data), "cuda")
    buf[_lib.HALO_LEFT:_lib.HALO_LEFT + len(data)] = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    comm.halo(ptr, len(data))
    cap = 1 << 16
    rec = torch.zeros((cap, 4), dtype=torch.int64, device="cuda")
    p = ag.Pattern("because each", k=2, linenum=1)
    res = comm.scan(p, ptr, len(data), 0, d_records=rec.data_ptr(), capacity=cap, ordinals=True)
    a = _oracle.compile("because each", k=2, linenum=1)
    cnt, recs = _oracle.scan(a, data)
    assert res.n_matched == cnt and [tuple(r[:3]) for r in rec[:res.n_records].cpu().tolist()] == list(recs)
    best, res = comm.bestmatch("goverment of the peple", ptr, len(data), 0, d_records=rec.data_ptr(), capacity=cap, nocase=1)
    best1, res1 = ag.bestmatch_device("goverment of the peple", ptr, len(data), nocase=1)
    assert (best, res.n_matched) == (best1, res1.n_matched) and res.n_records == res.n_matched

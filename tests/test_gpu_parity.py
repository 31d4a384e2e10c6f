"""GPU parity: the CUDA path (through the C ABI, host-buffer entry agb_scan_host) against the oracle on the
same inputs, and against the committed golden vectors of the real reference.  Bit-exact: counts and the
ordered (lasti, print_end) lists."""
import json, os, random
import pytest
import _oracle, _corpus
import agrep_b200 as ag
from agrep_b200 import _lib
from test_oracle_golden import G, CORPORA, text

pytestmark = pytest.mark.gpu


def api_kw(okw):
    return {k: v for k, v in okw.items() if k != "width"}


def check(pattern, data, **kw):
    a = _oracle.compile(pattern, **kw)
    cnt, recs = _oracle.scan(a, data)
    p = ag.Pattern(pattern, **api_kw(kw))
    res, got = p.scan_host(data)
    assert res.n_matched == cnt, (pattern, kw, res.n_matched, cnt)
    assert [(b, e) for b, e, _, _ in got] == [(b, e) for b, e, _ in recs], (pattern, kw)
    res2, _ = p.scan_host(data, want_records=False)
    assert res2.n_matched == cnt
    return res


@pytest.mark.parametrize("name", sorted(G["scan"]))
def test_golden_case(name):
    c = G["scan"][name]
    kw = dict(c["api"])
    res = check(c["pattern"], text(c["corpus"]), **kw)
    assert res.n_matched == c["count"]          # the real reference's answer


TEXT = _corpus.make_text(4000, seed=12345)


def test_random_differential():
    rnd = random.Random(77)
    lines = TEXT.decode().split("\n")
    n = 0
    while n < 60:
        ln = rnd.choice([l for l in lines if len(l) > 40])
        m = rnd.choice([3, 4, 5, 6, 8, 12, 16, 20, 24, 27, 33, 40])
        if m >= len(ln):
            continue
        st = rnd.randrange(len(ln) - m)
        pat = _corpus.mutate(rnd, ln[st:st + m], rnd.randint(0, 2))
        if any(ch in pat for ch in ";,.*-[]()<>|#{}~^$\\"):
            continue
        k = rnd.randint(0, min(4, len(pat) - 1))
        kw = dict(k=k, linenum=1)
        if rnd.random() < 0.3: kw["nocase"] = 1
        if rnd.random() < 0.2: kw["wordbound"] = 1
        check(pat, TEXT, **kw)
        n += 1


@pytest.mark.parametrize("data", [
    b"", b"\n", b"a", b"abc", b"abc\n", b"\nabc", b"\n\n\n", b"xabcx", b"ab\ncab\nc\n", b"abc\nabc\nabc",
    b"ab" + b"c" * 15 + b"\n", b"x" * 15 + b"abc\n", b"x" * 14 + b"abc\n" + b"y" * 40 + b"\nabc",
    b"x" * 5000 + b"abc" + b"y" * 5000 + b"\nzzz\n", b"\n" * 100 + b"abc" + b"\n" * 100])
@pytest.mark.parametrize("kw", [dict(k=0, linenum=1), dict(k=1, linenum=1), dict(k=0), dict(k=0, linenum=1, inverse=1),
                                dict(k=2, linenum=1, wordbound=1)])
def test_edges(data, kw):
    check("abc", data, **kw)


@pytest.mark.parametrize("delim,kw", [("$$", dict(k=1)), ("$$", dict(k=0, wordbound=1)), ("the", dict(k=1)),
                                       ("ab", dict(k=0)), ("\\.", dict(k=2)), ("$$", dict(k=2, inverse=1))])
def test_user_delimiters(delim, kw):
    para = _corpus.make_text(1500, seed=5, paragraphs=True)
    for data in (para, para + b"\n", b"\n\n" + para, para[:-1], b"\n\n\n\n\n" + para + b"\n\n\n"):
        check("because", data, linenum=1, delim=delim, **kw)
        check("state good", data, linenum=1, delim=delim, **kw)


def test_long_records_and_no_delimiter():
    rnd = random.Random(3)
    body = bytes(rnd.choice(b"abcdefghij ") for _ in range(300000))
    for data in (body, body + b"\n", body[:150000] + b"\n" + body[150000:], body.replace(b"j", b"\n")):
        check("abcde", data, k=1, linenum=1)
        check("fgh", data, k=0)
        check("a[bc]d", data, k=0, linenum=1)       # no anchors: every chunk goes to the record stage


def test_dense_forms_on_ragged_records():
    """no anchor plan (classes, -v, tiny patterns, `;`): the record stage walks every byte -- the slices form (fixed
    256-byte slices per thread, records spanning slices and 32 KiB tiles resolved afterwards) or, for '#', -p and
    run delimiters, the dense tile form.  Records of every awkward length: empty, 1, around the slice size, longer
    than a tile."""
    rnd = random.Random(11)
    words = [w for w in TEXT.decode().split() if w.isalpha()]
    lines = []
    for i in range(2500):
        ln = rnd.choice([0, 0, 1, 2, 5, 30, 60, 90, 255, 256, 257, 300, 511, 513, 1000, 4000])
        if i in (700, 1900): ln = 40000 + i
        row = ""
        while len(row) < ln: row += rnd.choice(words) + " "
        lines.append(row[:ln])
    data = "\n".join(lines).encode()
    for tail in (b"", b"\n", b"\n\n"):
        d = data + tail
        check("t[hx]e", d, k=0, linenum=1)
        check("th", d, k=1, linenum=1)
        check("because", d, k=2, linenum=1, inverse=1)
        check("people;state", d, k=1, linenum=1)
        check("gover#ent", d, k=1, linenum=1)
        check("wh[aeiou]ch", d, k=1, linenum=1, delim="e ")
        check("governmental", d, k=8, linenum=1)
    a = _oracle.compile("t[hx]ese", k=3, linenum=1)
    cnt, hist, recs = _oracle.scan_levels(a, 3, data)
    res, got = ag.Pattern("t[hx]ese", k=3, linenum=1).scan_host(data, levels=True)
    assert res.n_matched == cnt and list(res.level_hist)[:4] == hist[:4]
    assert [(b, e, l) for b, e, _, l in got] == [(b, e, l) for b, e, _, l in recs]


@pytest.mark.parametrize("pattern,kw,corpus_kw", [
    ("because each", dict(k=2, linenum=1), dict(nlines=2000, seed=31)),
    ("the", dict(k=0, linenum=1, wordbound=1), dict(nlines=800, seed=32, trailing_newline=False)),
    ("state", dict(k=1, linenum=1, delim="$$"), dict(nlines=1500, seed=33, paragraphs=True)),
    ("world", dict(k=1, linenum=1, delim="the"), dict(nlines=600, seed=34)),
    ("governmental", dict(k=5, linenum=1), dict(nlines=1500, seed=35)),
    ("t[hx]e", dict(k=0, linenum=1), dict(nlines=3000, seed=36)),
])
def test_device_ordinals_reproduce_j(pattern, kw, corpus_kw):
    """AGB_WANT_ORDINALS: the j that -n prints, counted on the device (k_delim_count + k_ordinals), against the
    oracle's j (pinned to the reference's -n output) and against the host helper; Result.n_closes = j at EOF"""
    data = _corpus.make_text(**corpus_kw)
    a = _oracle.compile(pattern, **kw)
    cnt, recs = _oracle.scan(a, data)
    assert cnt > 0
    p = ag.Pattern(pattern, **kw)
    for d in (data, b"\n" + data, data + b"\n\n\n", (kw.get("delim", "") .replace("$$", "\n\n")).encode() + data):
        cnt, recs = _oracle.scan(a, d)
        res, got = p.scan_host(d, ordinals=True)
        assert [(b, e, j) for b, e, j, _ in got] == [(b, e, j) for b, e, j in recs], (pattern, kw)
        arr = (_lib.Record * max(cnt, 1))()
        for i, (b, e, j) in enumerate(recs):
            arr[i].begin, arr[i].end = b, e
        _lib.lib().agb_fill_ordinals(p._h, d, len(d), arr, cnt)
        assert [arr[i].ordinal for i in range(cnt)] == [j for _, _, j in recs]
        # j at EOF: what a following shard adds to its ordinals (shard.ordinal_base)
        from agrep_b200 import shard
        assert res.n_closes == shard.count_closes(d, kw.get("delim", "\n").replace("$$", "\n\n").encode())


def test_random_metachar_differential():
    """random patterns with classes, '.', '#', <>, ',' and ';', anchors, under -i/-w/-v/-p/-x/-S2 and user delimiters:
    the list with ordinals and the count-only call against the oracle (tools/gpu_fuzz.py is the longer version)."""
    base = _corpus.make_text(3000, seed=5)
    words = [w for w in base.decode().split() if w.isalpha()]
    rnd = random.Random(404)

    def rand_pattern():
        w = (rnd.choice(words) + " " + rnd.choice(words))[:rnd.randint(3, 20)]
        out = []
        for ch in w:
            r = rnd.random()
            out.append("." if r < 0.08 else "[" + ch + "x]" if r < 0.12 else "[^q]" if r < 0.15 else "#" if r < 0.17
                       else ch.upper() if r < 0.19 else ch)
        p = "".join(out)
        r = rnd.random()
        return ("<" + p[:2] + ">" + p[2:] if r < 0.08 else p + "," + rnd.choice(words) if r < 0.14
                else p + ";" + rnd.choice(words) if r < 0.20 else "^" + p if r < 0.24 else p + "$" if r < 0.28 else p)
    done = 0
    for _ in range(220):
        data = ("\n".join(base.decode().split("\n")[:rnd.randint(1000, 2900)]) + rnd.choice(["\n", "", "\n\n"])).encode()
        pat = rand_pattern()
        kw = dict(k=rnd.choice([0, 0, 1, 2, 3, 4, 6, 8]))
        if rnd.random() < 0.8: kw["linenum"] = 1
        for p_, key in ((0.25, "nocase"), (0.15, "wordbound"), (0.1, "inverse"), (0.05, "ins_free"), (0.04, "wholeline")):
            if rnd.random() < p_: kw[key] = 1
        if rnd.random() < 0.15: kw["delim"] = rnd.choice(["$$", "e ", "ab", "\\.", "e e", "ee "])
        if kw["k"] and rnd.random() < 0.06: kw["cost_s"] = 2
        try:
            a = _oracle.compile(pat, **kw)
        except _oracle.OracleError:
            continue
        try:
            p = ag.Pattern(pat, **api_kw(kw))
        except ag.AgrepError as e:
            assert "delimiter" in str(e) and kw.get("ins_free"), (pat, kw, e)     # the documented refusal: -p with a multi-byte delimiter
            continue
        cnt, recs = _oracle.scan(a, data)
        res, got = p.scan_host(data, ordinals=True)
        keep = (lambda t: t[:3]) if a.engine != 4 else (lambda t: t[:2])      # sgrep/bm has no j
        assert res.n_matched == cnt and [keep(t) for t in got] == [keep(t) for t in recs], (pat, kw)
        assert p.scan_host(data, want_records=False)[0].n_matched == cnt, (pat, kw)
        done += 1
    assert done > 120


def test_wide_pattern_64bit_rows():
    pat = "people how too little state good very make"      # 42 chars -> M = 44
    data = TEXT + b"xx people how too little state good very make yy\nxx people hxw too litle state good very make\n"
    for k in (0, 1, 2, 3):
        res = check(pat, data, k=k, linenum=1)
    assert res.n_matched >= 2


def test_costs_and_insfree():
    check("between both life", TEXT, k=3, linenum=1, cost_s=2)
    check("between both life", TEXT, k=3, linenum=1, cost_i=2, cost_d=3)
    check("government", TEXT, k=2, linenum=1, ins_free=1)
    check("gover#ment;world", TEXT, k=1, linenum=1)


def test_levels_histogram():
    a = _oracle.compile("governmental", k=4, linenum=1)
    cnt, hist, recs = _oracle.scan_levels(a, 4, TEXT)
    p = ag.Pattern("governmental", k=4, linenum=1)
    res, got = p.scan_host(TEXT, levels=True)
    assert list(res.level_hist)[:5] == hist[:5]
    assert res.n_matched == cnt
    assert [(b, e, l) for b, e, _, l in got] == [(b, e, l) for b, e, _, l in recs]


def test_device_corpus_equals_host_corpus_and_device_scan():
    import torch
    n = 256 * 4096
    t = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    ag.corpus_device(t.data_ptr(), n, needle="because each", needle_every=8, needle_maxedits=3)
    torch.cuda.synchronize()
    host = ag.corpus_host(n, needle="because each", needle_every=8, needle_maxedits=3)
    assert bytes(t[:n].cpu().numpy().tobytes()) == host
    for k in (0, 1, 2, 3):
        a = _oracle.compile("because each", k=k, linenum=1)
        cnt, _ = _oracle.scan(a, host, want_records=False)
        p = ag.Pattern("because each", k=k, linenum=1)
        res = p.scan_device(t.data_ptr(), n)
        assert res.n_matched == cnt
        assert res.n_flagged < n // 16 // 20     # the anchor filter is selective on this text


def test_bestmatch_sweep():
    import torch
    host = ag.corpus_host(64 * 4096)
    t = torch.frombuffer(bytearray(host + b"\0" * 64), dtype=torch.uint8).cuda()
    for pat in ("goverment of the peple", "because each", "zzzzqqqqxxxx"):
        a0 = _oracle.compile(pat, k=0, linenum=1, nocase=1)
        want = -1
        for k in range(0, min(8, len(pat) - 1) + 1):
            a = _oracle.compile(pat, k=k, linenum=1, nocase=1)
            cnt, _ = _oracle.scan(a, host, want_records=False)
            if cnt:
                want = (k, cnt)
                break
        best, res = ag.bestmatch_device(pat, t.data_ptr(), len(host), nocase=1)
        if want == -1:
            assert best == -1
        else:
            assert (best, res.n_matched) == want


def test_latin1_case_folding_of_the_exact_engine():
    """-i with k=0 goes through LUT[] = CP[ISO-8859-1].lower_1 (bitap.c:171): 'É' (0xC9) matches 'é' (0xE9);
    with k>0 asearch() applies no LUT (asearch.c:96), so it does not."""
    body = (b"xx caf\xe9 noir yy\nxx CAF\xc9 NOIR yy\nxx cafe noir\nzz \x80\x87 qq caf\xc9\n" * 50) + TEXT[:20000]
    for kw in (dict(k=0, linenum=1, nocase=1), dict(k=1, linenum=1, nocase=1), dict(k=0, linenum=1), dict(k=0, nocase=1)):
        check(b"caf\xe9 noir", body, **kw)
        check(b"\xf6l qq", body.replace(b"cafe noir", b"\xd6L QQ \xf6l qq"), **kw)
    res = check(b"caf\xe9 noir", body, k=0, linenum=1, nocase=1)
    assert res.n_matched == 100


def test_host_entry_points_agree(tmp_path):
    """agb_scan_host on pageable and on page-locked memory, and agb_scan_fd on a regular file and on a pipe
    (the fill_buf replacement, bitap.c:450-477), all give the device-resident answer -- across several 64 MiB slices."""
    import ctypes, os, torch
    from agrep_b200 import _lib
    L = _lib.lib()
    n = (160 << 20) + 12345                      # 3 slices, ragged tail
    host = ag.corpus_host((n + 4095) // 4096 * 4096, needle="because each", needle_every=512, needle_maxedits=3)[:n]
    p = ag.Pattern("because each", k=2)
    dev = torch.frombuffer(bytearray(host + b"\0" * 64), dtype=torch.uint8).cuda()
    cap = 1 << 16
    drec = torch.zeros((cap, 4), dtype=torch.int64, device="cuda")
    ref = p.scan_device(dev.data_ptr(), n, d_records=drec.data_ptr(), capacity=cap)
    want = drec[:ref.n_records, :2].cpu().tolist()
    assert ref.n_matched > 100

    def via(fn):
        recs = (_lib.Record * cap)()
        res = _lib.Result()
        rc = fn(recs, res)
        assert rc == 0, L.agb_last_error()
        assert res.n_matched == ref.n_matched
        assert [[recs[i].begin, recs[i].end] for i in range(res.n_records)] == want

    buf = ctypes.create_string_buffer(host, n)
    via(lambda recs, res: L.agb_scan_host(p._h, buf, n, _lib.WANT_RECORDS, recs, cap, ctypes.byref(res)))          # pageable
    pinned = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    pinned.copy_(torch.frombuffer(bytearray(host), dtype=torch.uint8))
    via(lambda recs, res: L.agb_scan_host(p._h, ctypes.c_void_p(pinned.data_ptr()), n, _lib.WANT_RECORDS, recs, cap, ctypes.byref(res)))
    path = tmp_path / "corpus.txt"
    path.write_bytes(host)
    fd = os.open(str(path), os.O_RDONLY)
    try:
        via(lambda recs, res: L.agb_scan_fd(p._h, fd, _lib.WANT_RECORDS, recs, cap, ctypes.byref(res)))               # regular file
        assert os.lseek(fd, 0, os.SEEK_CUR) == n                      # the descriptor is left at the end, as read(2) would
        # the same past the page cache (AGB_ODIRECT=1: a second descriptor with O_DIRECT, whole 4 KiB blocks into the
        # page-aligned ring, the ragged last block short; silently the plain path where the file system refuses), and
        # from an offset that is not block aligned (no direct reads then)
        os.environ["AGB_ODIRECT"] = "1"
        try:
            os.lseek(fd, 0, os.SEEK_SET)
            via(lambda recs, res: L.agb_scan_fd(p._h, fd, _lib.WANT_RECORDS, recs, cap, ctypes.byref(res)))
            t = ctypes.c_void_p()
            os.lseek(fd, 0, os.SEEK_SET)
            assert L.agb_text_from_fd(fd, ctypes.byref(t)) == 0, L.agb_last_error()
            via(lambda recs, res: L.agb_scan_text(p._h, t, _lib.WANT_RECORDS, recs, cap, ctypes.byref(res)))
            L.agb_text_free(t)
        finally:
            del os.environ["AGB_ODIRECT"]
    finally:
        os.close(fd)
    r, w = os.pipe()
    small = host[:3 << 20]
    pid = os.fork()
    if pid == 0:
        os.close(r)
        os.write(w, small) if False else [os.write(w, small[i:i + 65536]) for i in range(0, len(small), 65536)]
        os._exit(0)
    os.close(w)
    recs = (_lib.Record * cap)()
    res = _lib.Result()
    assert L.agb_scan_fd(p._h, r, _lib.WANT_RECORDS, recs, cap, ctypes.byref(res)) == 0                               # pipe
    os.close(r); os.waitpid(pid, 0)
    a = _oracle.compile("because each", k=2, linenum=1)
    cnt, orecs = _oracle.scan(a, small)
    assert res.n_matched == cnt and [(recs[i].begin, recs[i].end) for i in range(res.n_records)] == [(b, e) for b, e, _ in orecs]


def test_exact_literal_count_path():
    """`agrep -c the`: an exact literal no longer than its anchor, count only, texts of 1 MiB and more -> the pass that
    needs no automaton (k_front_exact + k_exact_reduce: records that hold a hit, stitched from chunk, warp and block
    summaries).  Against the oracle, and against the same library's list path; ragged ends: no final newline, the text
    ending inside an occurrence, an occurrence at the very end."""
    body = _corpus.make_text(60000, seed=77)
    assert len(body) > (2 << 20)
    cases = [("the", {}), ("The", {}), ("and", {}), ("that", {}), ("of", {}), ("the", dict(linenum=1)), ("the", dict(linenum=1, nocase=1)),
             ("state", {}), ("e", dict(linenum=1)), ("the", dict(delim=";")), ("zqx", {})]
    texts = [body, body[:-1], body + b"xx th", body + b"the", body.replace(b"\n", b";", 20000), b"the" * 400000, body.replace(b" ", b"\n")]
    for data in texts:
        for pat, kw in cases:
            try:
                a = _oracle.compile(pat, **kw)
            except _oracle.OracleError:
                continue
            cnt, _ = _oracle.scan(a, data, want_records=False)
            p = ag.Pattern(pat, **api_kw(kw))
            got = p.scan_host(data, want_records=False)[0].n_matched          # host entry (streamed: the automaton forms)
            import torch
            t = torch.frombuffer(bytearray(data + b"\0" * 64), dtype=torch.uint8).cuda()
            dev = p.scan_device(t.data_ptr(), len(data)).n_matched             # device entry, count only: the exact pass when it applies
            assert got == cnt and dev == cnt, (pat, kw, len(data), got, dev, cnt)


@pytest.mark.parametrize("delim", ["X", "ab", "e ", "xx", "Q\\."])
def test_case_insensitive_delimiters(delim):
    """-i lower-cases the whole internal pattern, the delimiter included (maskgen.c:52-58, 259-266): -i -d X ends records
    at 'x' and at 'X'.  The automaton does that by itself; the code that finds delimiters by their bytes (record starts,
    ordinals, the delimiter counts of stage 1) goes by agb_desc.delim_fold."""
    rnd = random.Random(9)
    base = _corpus.make_text(2500, seed=61).decode()
    raw = delim.replace("\\", "")
    out = []
    for ln in base.split("\n"):
        out.append(ln)
        out.append(rnd.choice([raw, raw.upper(), raw.lower(), raw.swapcase(), raw + raw.upper()]))
    data = "".join(out).encode()
    for pat, kw in (("because", dict(k=1, linenum=1)), ("state good", dict(k=2, linenum=1)), ("the", dict(k=0, linenum=1)),
                    ("people", dict(k=0, linenum=1, inverse=1)), ("t[hx]e", dict(k=1, linenum=1))):
        kw = dict(kw, nocase=1, delim=delim)
        a = _oracle.compile(pat, **kw)
        for d in (data, raw.upper().encode() + data, data + raw.lower().encode()):
            cnt, recs = _oracle.scan(a, d)
            res, got = ag.Pattern(pat, **kw).scan_host(d, ordinals=True)
            assert res.n_matched == cnt and cnt > 0, (pat, delim)
            assert [t[:3] for t in got] == list(recs), (pat, delim)


@pytest.mark.parametrize("delim", ["aba", "abab", "=-=", "e e", "xyx"])
def test_self_overlapping_delimiters(delim):
    """delimiters that overlap themselves and are not runs (delim_kind 2): occurrences are taken from the left, one that shares
    a byte with the one taken before it is dropped; the record stages find record starts by walking the chain of overlapping
    occurrences back to its first.  List with ordinals and count against the oracle, in the list form (anchors), the tile
    form (classes, -v) and with errors."""
    from _corpus import overlap_text
    for seed in (3, 4, 5):
        data = overlap_text(delim, seed)
        for d in (data, delim.encode() + data, data + delim.encode(), data[:-len(delim)] + delim.encode()[:-1]):
            for pat, kw in (("state", dict(k=1, linenum=1)), ("e", dict(k=0, linenum=1)), ("world", dict(k=0, linenum=1, inverse=1)),
                            ("[st]tat.", dict(k=0, linenum=1)), ("because", dict(k=2, linenum=1, nocase=1))):
                a = _oracle.compile(pat, delim=delim, **kw)
                cnt, recs = _oracle.scan(a, d)
                p = ag.Pattern(pat, delim=delim, **api_kw(kw))
                res, got = p.scan_host(d, ordinals=True)
                assert res.n_matched == cnt and [t[:3] for t in got] == [t[:3] for t in recs], (delim, pat, seed)
                assert p.scan_host(d, want_records=False)[0].n_matched == cnt


def test_class_variant_anchor_plans():
    """positions that accept two bytes ([ea]) may sit inside an anchor: the piece "b[ea]c" stands in stage 1 as the two
    anchors "bec" and "bac" at the same place (pattern.c collect_runs) -- patterns with small classes keep the filter path
    instead of walking every byte.  Lists with ordinals and counts against the oracle, both variants present in the text."""
    data = _corpus.make_text(4000, seed=21)
    for a, b_, cnt in ((b"because", b"bacause", 60), (b"state", b"stote", 40), (b"government", b"govarnmant", 30), (b"government", b"governmant", 30),
                       (b"national", b"notional", 40), (b"order", b"ordar", 40), (b"world", b"warld", 40), (b"people", b"paopla", 40)):
        parts = data.split(a)
        step = max(1, len(parts) // cnt)
        data = b"".join(p + (b_ if i % step == 0 else a) for i, p in enumerate(parts[:-1])) + parts[-1]
    chunks = len(data) // 16
    for pat, kw, min_na in (("b[ea]cause", dict(k=1, linenum=1), 3), ("b[ea]cause", dict(k=0, linenum=1), 1), ("st[ao]te", dict(k=0, linenum=1), 2),
                            ("gov[ea]rnm[ea]nt", dict(k=1, linenum=1), 4),
                            ("w[oa]rld", dict(k=0, linenum=1, nocase=1), 2), ("p[ea]opl[ea] how", dict(k=2, linenum=1), 4),
                            ("b[ea]c.u[s-t]e", dict(k=1, linenum=1), 4), ("st[ao]te", dict(k=1, linenum=1, wordbound=1), 2)):
        p = ag.Pattern(pat, **api_kw(kw))
        assert p.desc.plan == 1 and p.desc.n_anchors >= min_na, (pat, p.desc.plan, p.desc.n_anchors)
        a = _oracle.compile(pat, **kw)
        cnt, recs = _oracle.scan(a, data)
        res, got = p.scan_host(data, ordinals=True)
        assert cnt > 5 and res.n_matched == cnt and [t[:3] for t in got] == [t[:3] for t in recs], (pat, kw)
        assert res.n_flagged < chunks // 4, (pat, res.n_flagged, chunks)
        assert p.scan_host(data, want_records=False)[0].n_matched == cnt, (pat, kw)


def test_inverse_count_by_complement():
    """`agrep -c -v`, newline records, device-resident text of 1 MiB and more: the number of records minus the number of
    matching records (scan.cu complement_usable) instead of the automaton over every byte.  Against the oracle; texts with
    and without a final newline, with blank lines, starting with a newline, ending in a match."""
    import torch
    body = _corpus.make_text(40000, seed=78)
    assert len(body) > (1 << 20) + 4096
    texts = [body, body[:-1], b"\n\n" + body, body.replace(b"the\n", b"the\n\n\n", 500), body + b"because each", body + b"\n\n\n",
             body + b"x because each y"]
    cases = [("because each", dict(k=2, inverse=1, linenum=1)), ("because each", dict(k=0, inverse=1, linenum=1)), ("state", dict(k=1, inverse=1, linenum=1)),
             ("government", dict(k=3, inverse=1, nocase=1)), ("b[ea]cause", dict(k=1, inverse=1, linenum=1)),
             ("world", dict(k=1, inverse=1, wordbound=1, linenum=1)), ("because each", dict(k=2, inverse=1, cost_s=2))]
    for data in texts:
        t = torch.frombuffer(bytearray(data + b"\0" * 64), dtype=torch.uint8).cuda()
        chunks = len(data) // 16
        for pat, kw in cases:
            try:
                a = _oracle.compile(pat, **kw)
            except _oracle.OracleError:
                continue
            cnt, _ = _oracle.scan(a, data, want_records=False)
            p = ag.Pattern(pat, **api_kw(kw))
            r = p.scan_device(t.data_ptr(), len(data))
            assert r.n_matched == cnt, (pat, kw, len(data), r.n_matched, cnt)
            assert r.n_flagged < chunks // 2, (pat, kw, r.n_flagged)              # the complement path ran (the every-byte forms flag all)
            assert p.scan_host(data, want_records=False)[0].n_matched == cnt       # the streaming entry: the automaton forms


def test_long_simple_literals_keep_sgrep_semantics():
    """a simple literal of more than 20 characters at k = 0: the reference runs monkey() instead of bm() (sgrep.c:407-442,
    1540-1834) with the same record semantics -- ASCII case folded, -w by isalnum neighbours (pinned on the reference binary in
    test_oracle_vs_reference.py::test_sgrep_long_literals_take_monkey); here the sgrep engine with 64-bit rows on the device."""
    pat = "homogeneous approximate matching"
    data = (TEXT[:200000] + b"xx Homogeneous Approximate Matching yy\nno homogeneous approximate matchin here\n"
            + b"ahomogeneous approximate matching\nthe homogeneous approximate matching.\n" + TEXT[200000:400000]
            + b"end homogeneous approximate matching")
    for kw in ({}, dict(wordbound=1), dict(delim=";")):
        a = _oracle.compile(pat, **kw)
        assert a.engine == 4
        d = data.replace(b"\n", b";") if kw.get("delim") else data
        cnt, recs = _oracle.scan(a, d)
        p = ag.Pattern(pat, **api_kw(kw))
        res, got = p.scan_host(d)
        assert cnt >= 3 and res.n_matched == cnt and [t[:2] for t in got] == [t[:2] for t in recs], (kw, cnt, res.n_matched)
        assert p.scan_host(d, want_records=False)[0].n_matched == cnt

"""Oracle vs the committed golden vectors (tests/golden/reference_vectors.json, produced from the
unmodified reference by tests/golden/make_golden.py).  Runs everywhere, no reference needed."""
import json, os
import pytest
import _oracle, _corpus

G = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_vectors.json")))
from golden.make_golden import CORPORA  # corpus recipes (seeded)

_texts = {}


def text(name):
    if name not in _texts:
        _texts[name] = _corpus.make_text(**CORPORA[name])
    return _texts[name]


@pytest.mark.parametrize("name", sorted(G["scan"]))
def test_scan_case(name):
    c = G["scan"][name]
    kw = dict(c["api"])
    bm = "linenum" not in kw
    a = _oracle.compile(c["pattern"], **({} if bm else {"width": 32}), **kw)
    cnt, recs = _oracle.scan(a, text(c["corpus"]))
    assert cnt == c["count"]
    if "ordinals" in c and "delim" not in kw:
        assert [r[2] - 1 for r in recs] == c["ordinals"]


def _args_to_kw(rargs):
    kw, it = dict(linenum=1, width=32), iter(rargs)
    for a in it:
        if a == "-n": pass
        elif a == "-i": kw["nocase"] = 1
        elif a == "-w": kw["wordbound"] = 1
        elif a == "-x": kw["wholeline"] = 1
        elif a == "-p": kw["ins_free"] = 1
        elif a == "-d": kw["delim"] = next(it)
        elif a[1:].isdigit(): kw["k"] = int(a[1:])
        elif a[1] == "I": kw["cost_i"] = int(a[2:])
        elif a[1] == "S": kw["cost_s"] = int(a[2:])
        elif a[1] == "D": kw["cost_d"] = int(a[2:])
        else: raise ValueError(a)
    return kw


@pytest.mark.parametrize("name", sorted(G["dump"]))
def test_automaton_words(name):
    d = G["dump"][name]
    a = _oracle.compile(d["pattern"], **_args_to_kw(d["ref_args"]))
    assert a.M == d["M"]
    m32 = 0xFFFFFFFF
    assert a.init0 & m32 == d["Init0"]
    # -p: Init1 is forced to all ones inside bitap()/asearch() (bitap.c:123), after maskgen; the dump is taken after the scan
    assert a.init1 & m32 == d["Init1"]
    assert a.noerr & m32 == d["NO_ERR_MASK"]
    assert a.endpos == d["endposition"]
    assert a.dendpos == d["D_endpos"]
    assert a.wildmask == d["wildmask"]
    assert a.and_mode == d["AND"]
    for c in range(256):
        assert a.mask[c] == d["mask"].get(str(c), 0), c


def test_lut_lower1():
    import ctypes
    lut = (ctypes.c_ubyte * 256)()
    _oracle.lib().orc_lut_lower1(lut)
    assert list(lut) == json.load(open(os.path.join(os.path.dirname(__file__), "golden", "lut_lower1.json")))

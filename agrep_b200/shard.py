"""Multi-GPU: the Python face of the C sharded scan (agb_scan_sharded, agrep_b200/csrc/shard.cu: the cut rule on the
device, ncclAllGather of headers and match lists) and, below it, a host restatement of the same cut rule that the CPU
tests use (gloo, world_size 2).

Host logic: shard a text by byte range at record boundaries, gather the per-shard match lists.

Records are independent once their boundaries are known (the automaton is reset at every delimiter,
asearch.c:175-196), so the scan path needs no data-path collective: rank r scans [cut[r], cut[r+1]) as a text
of its own and only the match lists travel.  The cut rule: the nominal cut i*n/world moves forward
to just after the first delimiter that ENDS at or after it; for run delimiters ($$ = "\\n\\n") the greedy,
non-overlapping parse is resolved from the start of the run, exactly as the device does (scan.cu delim_ends_at).
A shard therefore starts where a record starts and its scan begins with the same virtual '\\n' the reference
puts in front of a file (bitap.c:140) -- which is also what precedes the record in the whole text whenever the
delimiter ends in '\\n'.  (For delimiters not ending in '\\n' the first record of a shard sees '\\n' instead of
the delimiter's last byte as its re-fed byte; -w patterns can tell the two apart, so such texts are cut only
where the re-fed byte is irrelevant -- see `cut_points(..., strict=True)`.)
"""
import torch


def _delim_ends_at(text, q, delim):
    """is q the last byte of a record-closing delimiter (same rule as the device, scan.cu)?"""
    L = len(delim)
    if L == 1:
        return text[q] == delim[0]
    if any(delim[:b] == delim[L - b:] for b in range(1, L)) and len(set(delim)) > 1:
        # overlaps itself and is not a run ("aba"): occurrences are taken from the left, one that shares a byte with the one
        # taken before it is dropped; what a chain of overlapping occurrences yields depends on where it starts
        occ = lambda e: e + 1 >= L and bytes(text[e + 1 - L:e + 1]) == bytes(delim)
        if not occ(q):
            return False
        e = q
        while True:
            prev = next((c for c in range(e - L + 1, e) if occ(c)), None)
            if prev is None:
                break
            e = prev
        last = e
        for c in range(e + 1, q + 1):
            if c - L + 1 > last and occ(c):
                last = c
        return last == q
    if any(delim[:b] == delim[L - b:] for b in range(1, L)):          # a run c^L: pairs from the start of the run
        c = delim[0]
        if text[q] != c:
            return False
        run, p = 1, q - 1
        while p >= 0 and text[p] == c:
            run += 1
            p -= 1
        if p < 0 and c == 0x0A:
            run += 1                                                    # the virtual '\n' in front of the text
        return run % L == 0
    return q + 1 >= L and bytes(text[q + 1 - L:q + 1]) == bytes(delim)


def cut_points(text, world, delim=b"\n"):
    """[0, c1, ..., n]: shard r is text[c[r]:c[r+1]]; every interior cut is the first byte after a delimiter."""
    n = len(text)
    cuts = [0]
    for r in range(1, world):
        q = max(cuts[-1], (n * r) // world)
        while q < n and not _delim_ends_at(text, q, delim):
            q += 1
        cuts.append(min(q + 1, n))
    cuts.append(n)
    return cuts


def page_shards(total_bytes, world, page=4096):
    """synthetic corpus: records never cross a 4 KiB page, so equal page-aligned ranges are record aligned"""
    per = total_bytes // (page * world) * page
    return [(r * per, per) for r in range(world)]


def count_closes(text, delim=b"\n"):
    """record closes of a text scanned on its own: delimiter ends at the virtual '\n', in the text and in the delimiter
    appended at EOF -- what agb_result.n_closes reports (host restatement for the CPU tests)"""
    L = len(delim)
    ext = bytes(text) + bytes(delim)
    n = sum(1 for q in range(len(ext)) if _delim_ends_at(ext, q, delim))
    return n + (1 if (L == 1 and delim[0] == 0x0A) else 0)


def ordinal_base(closes_before, rank, delim=b"\n"):
    """what rank `rank` adds to its shard-local ordinals (the j that -n prints): the closes of the shards before it,
    minus what those scans counted that the whole text does not have -- the delimiter appended at each shard's EOF
    and, for the 1-byte '\n', this and every other later shard's virtual '\n' (it IS the '\n' that ended the
    shard before)."""
    virt = 1 if (len(delim) == 1 and delim[0] == 0x0A) else 0
    return sum(closes_before) - rank * (1 + virt)


def gather_records(recs, n_records, base, dist=None, group=None, closes=None, delim=b"\n", shard_head=None):
    """recs: int64 tensor [cap, 4] of (begin, end, ordinal, level) with shard-local offsets, n_records valid rows.
    Returns on every rank the concatenation over ranks, offsets made global (+ base) -- ordered because shards are.
    closes: this shard's agb_result.n_closes; when given the ordinals are made global too (an exclusive
    prefix sum over the ranks, riding on the same all_gather as the counts).
    Collectives: all_gather of the counts, all_gather of the lists padded to the longest (payload = 32 B/record)."""
    blk = recs[:n_records].clone()
    if n_records:
        blk[:, 0:2] += base
        # a scan reports begin = 0 for its first record when the text has a user delimiter that has not been seen yet; in
        # the whole text that record was opened by the delimiter that ends just before the shard (begin = base - L) --
        # unless the shard itself starts with a delimiter, then begin = 0 is that one (the first, empty record is never reported)
        if base > 0 and bytes(delim) != b"\n" and int(recs[0, 0]) == 0 and not (shard_head is not None and bytes(shard_head[:len(delim)]) == bytes(delim)):
            blk[0, 0] = base - len(delim)
    if dist is None or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return blk
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    cnt = torch.tensor([n_records, closes if closes is not None else 0], dtype=torch.int64, device=recs.device)
    allc = torch.zeros(2 * world, dtype=torch.int64, device=recs.device)
    dist.all_gather_into_tensor(allc, cnt, group=group)
    pairs = allc.view(world, 2).tolist()
    counts = [int(x[0]) for x in pairs]
    if closes is not None and n_records:
        blk[:, 2] += ordinal_base([int(x[1]) for x in pairs[:rank]], rank, delim)
    m = max(counts)
    if m == 0:
        return blk
    pad = torch.zeros((m, 4), dtype=torch.int64, device=recs.device)
    pad[:n_records] = blk
    out = torch.empty((world * m, 4), dtype=torch.int64, device=recs.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    return torch.cat([out[r * m:r * m + counts[r]] for r in range(world)], dim=0)


# ---- the C path: one rank per GPU, NCCL inside libagrepb200.so -----------------------------------------------------
import ctypes as _C
from . import _lib as _L


class Comm:
    """agb_comm: an NCCL communicator owned by the library.  The 128-byte unique id is made on rank 0 and handed to the
    other ranks through whatever the caller already has (here: torch.distributed, any backend)."""

    def __init__(self, dist=None, world=None, rank=None, unique_id=None):
        lib = _L.lib()
        if dist is not None and dist.is_initialized():
            world, rank = dist.get_world_size(), dist.get_rank()
            box = [None]
            if rank == 0:
                buf = (_C.c_ubyte * 128)()
                if lib.agb_comm_unique_id(buf) != 0:
                    raise RuntimeError(lib.agb_last_error().decode())
                box[0] = bytes(buf)
            dist.broadcast_object_list(box, src=0)
            unique_id = box[0]
        elif unique_id is None:
            world, rank = 1, 0
            buf = (_C.c_ubyte * 128)()
            if lib.agb_comm_unique_id(buf) != 0:
                raise RuntimeError(lib.agb_last_error().decode())
            unique_id = bytes(buf)
        self.world, self.rank = world, rank
        self._h = _C.c_void_p()
        idbuf = (_C.c_ubyte * 128).from_buffer_copy(unique_id)
        if lib.agb_comm_init(_C.byref(self._h), world, rank, idbuf) != 0:
            raise RuntimeError(lib.agb_last_error().decode())

    def halo(self, shard_ptr, n_local, stream=0):
        """fetch the halos of this rank's shard from its neighbours (once per text)"""
        if _L.lib().agb_shard_halo(self._h, _C.c_void_p(shard_ptr), n_local, _C.c_void_p(stream)) != 0:
            raise RuntimeError(_L.lib().agb_last_error().decode())

    def scan(self, pattern, shard_ptr, n_local, global_offset, d_records=0, capacity=0, stream=0, ordinals=False, levels=False):
        want = (_L.WANT_RECORDS if capacity else _L.WANT_COUNT) | (_L.WANT_ORDINALS if ordinals else 0) | (_L.WANT_LEVELS if levels else 0)
        res = _L.Result()
        rc = _L.lib().agb_scan_sharded(pattern._h, self._h, _C.c_void_p(shard_ptr), n_local, global_offset, want,
                                       _C.c_void_p(d_records), capacity, _C.c_void_p(stream), _C.byref(res))
        if rc != 0:
            raise RuntimeError("agb_scan_sharded rc=%d: %s" % (rc, _L.lib().agb_last_error().decode()))
        return res

    def bestmatch(self, pattern, shard_ptr, n_local, global_offset, d_records=0, capacity=0, stream=0, **kw):
        if isinstance(pattern, str):
            pattern = pattern.encode("latin-1")
        d = kw.pop("delim", None)
        if isinstance(d, str):
            d = d.encode("latin-1")
        o = _L.Options(delim=d, **{k: int(v) for k, v in kw.items()})
        res, best, err = _L.Result(), _C.c_int(-1), _C.create_string_buffer(512)
        rc = _L.lib().agb_bestmatch_sharded(pattern, _C.byref(o), self._h, _C.c_void_p(shard_ptr), n_local, global_offset,
                                            _C.c_void_p(d_records), capacity, _C.c_void_p(stream), _C.byref(best), _C.byref(res), err, 512)
        if rc != 0:
            raise RuntimeError(err.value.decode() or _L.lib().agb_last_error().decode())
        return best.value, res

    def __del__(self):
        try:
            if self._h:
                _L.lib().agb_comm_free(self._h)
        except Exception:
            pass


def shard_buffer(torch, n_local, device):
    """a device buffer with room for both halos around a shard of n_local bytes: (tensor, pointer of the shard itself)"""
    t = torch.zeros(_L.HALO_LEFT + n_local + _L.HALO_RIGHT + 64, dtype=torch.uint8, device=device)
    return t, t.data_ptr() + _L.HALO_LEFT

"""Worker of tests/test_gpu_shard_nccl.py: one process per GPU (torchrun), the text cut at multiples of 512 bytes -- in the
middle of records --, agb_shard_halo + agb_scan_sharded / agb_bestmatch_sharded over NCCL, and on every rank the gathered
answer against the oracle on the whole text."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import torch.distributed as dist
import _oracle, _corpus
import agrep_b200 as ag
from agrep_b200 import shard, _lib


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")                      # only to hand the NCCL unique id around; the data path is the library's
    comm = shard.Comm(dist)
    cases = [
        (_corpus.make_text(20000, seed=41), "because each", dict(k=2, linenum=1), b"\n"),
        (_corpus.make_text(20000, seed=42), "the", dict(), b"\n"),
        (_corpus.make_text(20000, seed=43, paragraphs=True), "state", dict(k=1, linenum=1, wordbound=1, delim="$$"), b"\n\n"),
        (_corpus.make_text(20000, seed=44), "people", dict(k=1, linenum=1, inverse=1), b"\n"),
        (_corpus.make_text(12000, seed=45).replace(b"\n", b"; "), "world", dict(k=1, linenum=1, delim="; "), b"; "),
    ]
    for data, pattern, kw, dbytes in cases:
        n = len(data)
        per = (n // world) // 512 * 512
        off = rank * per
        n_local = per if rank + 1 < world else n - off
        buf, ptr = shard.shard_buffer(torch, n_local, "cuda")
        buf[_lib.HALO_LEFT:_lib.HALO_LEFT + n_local] = torch.frombuffer(bytearray(data[off:off + n_local]), dtype=torch.uint8).cuda()
        comm.halo(ptr, n_local)
        cap = n + 2
        rec = torch.zeros((cap, 4), dtype=torch.int64, device="cuda")
        p = ag.Pattern(pattern, **kw)
        res = comm.scan(p, ptr, n_local, off, d_records=rec.data_ptr(), capacity=cap, ordinals=True)
        a = _oracle.compile(pattern, **kw)
        cnt, recs = _oracle.scan(a, data)
        got = [tuple(r[:3]) for r in rec[:res.n_records].cpu().tolist()]
        keep = (lambda t: t) if a.engine != 4 else (lambda t: t[:2])
        assert res.n_matched == cnt and cnt > 0, (rank, pattern, res.n_matched, cnt)
        assert [keep(t) for t in got] == [keep(t) for t in recs], (rank, pattern)
        assert res.n_closes == shard.count_closes(data, dbytes), (rank, pattern, res.n_closes)
        res2 = comm.scan(p, ptr, n_local, off)          # count only
        assert res2.n_matched == cnt
    # best match over the shards: the histogram is summed over the ranks, the list is the best level's of the whole text
    data = _corpus.make_text(30000, seed=46)
    n = len(data); per = (n // world) // 512 * 512; off = rank * per; n_local = per if rank + 1 < world else n - off
    buf, ptr = shard.shard_buffer(torch, n_local, "cuda")
    buf[_lib.HALO_LEFT:_lib.HALO_LEFT + n_local] = torch.frombuffer(bytearray(data[off:off + n_local]), dtype=torch.uint8).cuda()
    comm.halo(ptr, n_local)
    rec = torch.zeros((n + 2, 4), dtype=torch.int64, device="cuda")
    for pat in ("goverment of the peple", "because each", "zzzzqqqqxxxx"):
        best, res = comm.bestmatch(pat, ptr, n_local, off, d_records=rec.data_ptr(), capacity=n + 2, nocase=1)
        want = None
        for k in range(0, 9):
            if k >= len(pat):
                break
            a = _oracle.compile(pat, k=k, linenum=1, nocase=1)
            cnt, hist, lrecs = _oracle.scan_levels(a, k, data, want_level=k)
            if cnt and hist[k]:
                want = (k, hist[k], [(b, e) for b, e, _, lv in lrecs if lv == k])
                break
        if want is None:
            assert best == -1, (pat, best)
        else:
            got = [tuple(r[:2]) for r in rec[:res.n_records].cpu().tolist()]
            assert (best, res.n_matched) == want[:2] and got == want[2], (rank, pat, best, res.n_matched, want[:2])
    dist.barrier()
    if rank == 0:
        print("shard_nccl_worker ok: world", world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

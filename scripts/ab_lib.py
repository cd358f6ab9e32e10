"""A/B of two builds of libcmx.so on the SAME box (box-to-box variance of the pool is several percent): the same
search through the bare C ABI with either library, per-slab times from CMX_DEBUG_SLABS.
    python scripts/ab_lib.py <rows> <nq> <k> [lib ...]        (default libs: lib/libcmx_r1.so = round 1, lib/libcmx.so)
Only entry points both builds export are used (create / add / search_mixed / last_stats / set_profiling)."""
import ctypes as C, json, os, sys, pathlib, subprocess
ROOT = pathlib.Path(__file__).resolve().parent.parent
LIBDIR = ROOT / "codemix-dense-retrieval_b200" / "lib"


def child(lib_path, rows, nq, k, reps):
    sys.path[:0] = [str(ROOT), str(ROOT / "codemix-dense-retrieval_b200")]
    import torch
    import bench
    from cmx._lib import SearchStats
    L = C.CDLL(lib_path)
    L.cmx_last_error.restype = C.c_char_p
    vp = C.c_void_p

    def check(rc):
        if rc:
            raise RuntimeError(L.cmx_last_error().decode())

    dev = torch.device("cuda", 0)
    d = 1024
    ix = vp()
    check(L.cmx_index_create(C.c_int(d), C.c_int(0), C.byref(ix)))
    check(L.cmx_index_reserve(ix, C.c_int64(rows)))

    def add(x):
        torch.cuda.synchronize()
        check(L.cmx_index_add(ix, vp(x.data_ptr()), C.c_int64(x.shape[0]), C.c_int(1)))

    bench.fill_rows(add, 0, rows, d, dev, rows)
    P, S = bench.make_queries(nq, d, dev)
    D = torch.empty((1, nq, k), dtype=torch.float32, device=dev)
    I = torch.empty((1, nq, k), dtype=torch.int64, device=dev)
    alphas = (C.c_double * 1)(0.5)
    check(L.cmx_set_profiling(C.c_int(1)))

    def step():
        check(L.cmx_search_mixed(ix, vp(P.data_ptr()), vp(S.data_ptr()), C.c_int64(nq), alphas, C.c_int(1), C.c_int(k),
                                 vp(D.data_ptr()), vp(I.data_ptr()), None, C.c_int(1), C.c_int64(0), C.c_int(0), None))

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sc = se = 0.0
    st = SearchStats()
    e0.record()
    for _ in range(reps):
        step()
        check(L.cmx_index_last_stats(ix, C.byref(st)))
        sc += st.score_ms; se += st.select_ms
    e1.record(); torch.cuda.synchronize()
    print("RESULT " + json.dumps({"lib": os.path.basename(lib_path), "rows": rows, "nq": nq, "k": k, "ms_per_step": e0.elapsed_time(e1) / reps,
                                  "score_ms": sc / reps, "select_ms": se / reps, "slabs": st.slabs,
                                  "checksum": float(D.double().sum()), "idsum": int(I.sum())}), flush=True)


if __name__ == "__main__":
    if sys.argv[1] == "--child":
        child(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]))
        sys.exit(0)
    rows, nq, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    libs = sys.argv[4:] or [str(LIBDIR / "libcmx_r1.so"), str(LIBDIR / "libcmx.so")]
    reps = 6
    for rnd in range(2):  # A B A B
        for lib in libs:
            env = dict(os.environ, CMX_DEBUG_SLABS="1")
            out = subprocess.run([sys.executable, __file__, "--child", lib, str(rows), str(nq), str(k), str(reps)], env=env,
                                 capture_output=True, text=True)
            res = [l for l in out.stdout.splitlines() if l.startswith("RESULT ")]
            slabs = [l for l in out.stderr.splitlines() if l.startswith("[cmx] slab")]
            print(res[-1] if res else "FAILED " + out.stderr[-800:], flush=True)
            nsl = json.loads(res[-1][7:])["slabs"] if res else 0
            for l in slabs[-nsl:]:
                print("   ", l, flush=True)

"""Generate golden fixtures by running the reference's OWN Python sources.  TEST INFRASTRUCTURE.

Runs only in the build container (needs /root/reference).  The reference package cannot be
imported whole (mpi4py / h5py / matplotlib / cupy are absent), so its unmodified
``pyLOM/vmmath/{maths,averaging,truncation,svd,stats}.py`` and ``pyLOM/POD/wrapper.py`` are
loaded by file path under a stub ``pyLOM.utils`` (numpy aliased as ``cp`` exactly like
``pyLOM/utils/gpu.py:66`` does without cupy; ``mpi_send/mpi_recv`` backed by per-pair queues;
one *thread* per simulated rank, each with its own module instances because
``MPI_RANK/MPI_SIZE`` are bound at import time, ``pyLOM/vmmath/svd.py:14``).

Usage:  python oracle/gen_golden.py            -> writes tests/golden/*.npz
"""
import importlib.util, os, queue, sys, threading, types
import numpy as np

REF = os.environ.get("PYLOM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import synth  # noqa: E402


class _World:
    def __init__(self, size):
        self.size = size
        self.q = {(s, d): queue.Queue() for s in range(size) for d in range(size)}
        self.red = threading.Barrier(size)
        self.slots = [None] * size


def _load_rank(world, rank):
    """Return a namespace with the reference functions bound to (rank, world.size)."""
    tag = f"pyLOMref_r{rank}_of{world.size}_{id(world)}"
    def mod(name, path=False):
        m = types.ModuleType(name)
        if path:
            m.__path__ = []
        sys.modules[name] = m
        return m
    pk = mod(tag, True)
    utils = mod(tag + ".utils", True)
    gpu = mod(tag + ".utils.gpu")
    cr = mod(tag + ".utils.cr")
    vm = mod(tag + ".vmmath", True)
    pod = mod(tag + ".POD", True)
    gpu.cp = np
    gpu.gpu_to_cpu = lambda x: x
    gpu.cpu_to_gpu = lambda x: x
    ident = lambda name: (lambda f: f)
    cr.cr_nvtx = ident
    cr.cr_start = lambda *a, **k: None
    cr.cr_stop = lambda *a, **k: None
    def send(obj, dest):
        world.q[(rank, dest)].put(np.array(obj, copy=True))
    def recv(source=0):
        return world.q[(source, rank)].get()
    def reduce(x, root=0, op="sum", all=False):
        world.slots[rank] = np.array(x, copy=True)
        world.red.wait()
        tot = sum(world.slots[1:], world.slots[0].copy())
        world.red.wait()
        return tot
    utils.cr_nvtx = ident
    utils.MPI_RANK, utils.MPI_SIZE = rank, world.size
    utils.mpi_send, utils.mpi_recv, utils.mpi_reduce = send, recv, reduce
    utils.gpu, utils.cr = gpu, cr
    utils.gpu_to_cpu, utils.cpu_to_gpu = gpu.gpu_to_cpu, gpu.cpu_to_gpu
    def load(sub, rel):
        spec = importlib.util.spec_from_file_location(f"{tag}.{sub}", os.path.join(REF, "pyLOM", rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = m
        spec.loader.exec_module(m)
        return m
    out = types.SimpleNamespace()
    for name in ("maths", "averaging", "truncation", "stats", "svd"):
        m = load(f"vmmath.{name}", f"vmmath/{name}.py")
        for k, v in vars(m).items():
            if callable(v) and not k.startswith("_"):
                setattr(vm, k, v)
        setattr(out, name, m)
    out.POD = load("POD.wrapper", "POD/wrapper.py")
    for k, v in vars(out.POD).items():
        if callable(v) and not k.startswith("_"):
            setattr(pod, k, v)
    utils.cr_start, utils.cr_stop = cr.cr_start, cr.cr_stop
    mod(tag + ".DMD", True)
    out.DMD = load("DMD.wrapper", "DMD/wrapper.py")
    return out


def run_ranks(fn, shards):
    """Run ``fn(ref_namespace, shard)`` on len(shards) simulated ranks; return list of results."""
    P = len(shards)
    world = _World(P)
    refs = [_load_rank(world, r) for r in range(P)]
    res, err = [None] * P, []
    def work(r):
        try:
            res[r] = fn(refs[r], shards[r])
        except Exception as e:  # pragma: no cover
            err.append((r, e))
    th = [threading.Thread(target=work, args=(r,)) for r in range(P)]
    [t.start() for t in th]
    [t.join() for t in th]
    if err:
        raise err[0][1]
    return res


def split_rows(A, P):
    sys.path.insert(0, HERE)
    from pod_oracle import worksplit
    return [A[slice(*worksplit(0, A.shape[0], r, P))] for r in range(P)]


CASES = [
    # name, m, n, kind, seed, ranks
    ("ex4x2", 4, 2, "example_svd", 0, (1,)),            # Examples/example_SVD.py:17
    ("rand_300x8", 300, 8, "rand", 11, (1, 2, 3, 4)),
    ("synth_700x24", 700, 24, "synth", 2021, (1, 2, 4)),
    ("cond_640x33", 640, 33, "cond1e9", 7, (1, 3, 8)),
    ("cyl_twin_400x151", 400, 151, "synth", 2021, (1, 2)),   # cfg1 twin (n=151)
]


def make_input(m, n, kind, seed):
    if kind == "example_svd":
        return np.array([[1, 2], [3, 4], [5, 6], [7, 8]], dtype=np.double, order="C")
    if kind == "rand":
        return synth.random_matrix(m, n, seed)
    if kind == "cond1e9":
        return synth.random_matrix(m, n, seed, cond=1e9)
    return synth.snapshots(m, n, seed)


# randomized SVD (pyLOM/vmmath/svd.py:120-144,254-273): one rank only -- the reference seeds numpy's GLOBAL generator
# (svd.py:131,264), which the simulated ranks (threads) would race on; multi-rank parity goes through the oracle.
RSVD_CASES = [
    # name, m, n, kind, seed(input), r, q, seed(sketch)
    ("rsvd_synth_900x40", 900, 40, "synth", 2021, 8, 2, 7),
    ("rsvd_rand_500x24", 500, 24, "rand", 5, 6, 3, 123),
    ("rsvd_synth_1200x96_q0", 1200, 96, "synth", 2022, 12, 0, 99),
]


def main_rsvd(outdir):
    for name, m, n, kind, seed, r, q, sk in RSVD_CASES:
        A = make_input(m, n, kind, seed)
        ref = _load_rank(_World(1), 0)
        Q, B = ref.svd.randomized_qr(A, r, q, seed=sk)
        U, S, V = ref.svd.randomized_svd(A, r, q, seed=sk)
        Up, Sp, Vp = ref.POD.run(A, remove_mean=True, randomized=True, r=r, q=q, seed=sk)
        np.random.seed(sk)
        omega = np.random.rand(n, r)
        # streaming variant: first half of the snapshots, then the second half (svd.py:176-226)
        n1 = n // 2
        Qs1, Bs1, Ys1 = ref.svd.init_qr_streaming(np.ascontiguousarray(A[:, :n1]), r, q, seed=sk)
        Qs2, Bs2, Ys2 = ref.svd.update_qr_streaming(np.ascontiguousarray(A[:, n1:]), Qs1, Bs1, Ys1.copy(), r, q)
        blob = {"A": A, "r": np.array(r), "q": np.array(q), "seed": np.array(sk), "omega": omega,
                "Q": Q, "B": B, "U": U, "S": S, "V": V, "pod_U": Up, "pod_S": Sp, "pod_V": Vp,
                "st_n1": np.array(n1), "st_Q1": Qs1, "st_B1": Bs1, "st_Y1": Ys1, "st_Q2": Qs2, "st_B2": Bs2, "st_Y2": Ys2}
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, S[:3], os.path.getsize(path) // 1024, "KiB")


# DMD on the POD basis (pyLOM/DMD/wrapper.py:50-146): travelling / decaying waves + noise, so that the eigenvalues are
# well separated conjugate pairs.
DMD_CASES = [
    # name, m, n, seed, r, ranks
    ("dmd_waves_600x48", 600, 48, 3, 8, (1, 2)),
    ("dmd_waves_1500x80_res", 1500, 80, 4, 1e-4, (1, 3)),
]


def main_dmd(outdir):
    for name, m, n, seed, r, ranks in DMD_CASES:
        X = synth.dmd_waves(m, n, seed)
        blob = {"X": X, "r": np.array(r), "dt": np.array(0.1)}
        for P in ranks:
            res = run_ranks(lambda ref, Xi: ref.DMD.run(Xi, r, remove_mean=True), split_rows(X, P))
            muR, muI, _, b = res[0]
            Phi = np.vstack([x[2] for x in res])
            blob[f"P{P}_muReal"], blob[f"P{P}_muImag"], blob[f"P{P}_Phi"], blob[f"P{P}_b"] = muR, muI, Phi, b
            if P == 1:
                ref = _load_rank(_World(1), 0)
                delta, omega = ref.DMD.frequency_damping(muR, muI, 0.1)
                t = np.arange(n, dtype=np.double)
                blob["delta"], blob["omega"] = delta, omega
                blob["X_DMD"] = ref.DMD.reconstruction_jovanovic(Phi, muR, muI, t, b)
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, blob["P1_muReal"][:4], blob["P1_muImag"][:4], os.path.getsize(path) // 1024, "KiB")


# complex128 and float32 tsqr_svd: the reference's Python tsqr_svd is dtype-generic (numpy qr / svd), so the same
# unmodified file yields the fixtures for the ztsqr_svd / stsqr_svd variants (pyLOM/vmmath/src/svd.c:955-1010, 529-563).
DTYPE_CASES = [
    # name, m, n, dtype, seed, ranks, decay (column j scaled by 10^(-decay j / n))
    ("ztsqr_svd_500x12", 500, 12, np.complex128, 5, (1, 2, 3), 3.0),
    ("stsqr_svd_800x20", 800, 20, np.float32, 6, (1, 2), 2.0),
]


def main_dtypes(outdir):
    outdir = os.path.join(outdir, "aux")
    os.makedirs(outdir, exist_ok=True)
    for name, m, n, dt, seed, ranks, decay in DTYPE_CASES:
        rng = np.random.default_rng(seed)
        A = rng.standard_normal((m, n))
        if dt == np.complex128:
            A = A + 1j * rng.standard_normal((m, n))
        A = np.ascontiguousarray((A * 10.0 ** (-decay * np.arange(n) / n)).astype(dt))
        blob = {"A": A}
        for P in ranks:
            r = run_ranks(lambda ref, Ai: ref.svd.tsqr_svd(np.ascontiguousarray(Ai)), split_rows(A, P))
            blob[f"P{P}_U"] = np.vstack([x[0] for x in r])
            blob[f"P{P}_S"], blob[f"P{P}_V"] = r[0][1], r[0][2]
            assert blob[f"P{P}_U"].dtype == dt and all(np.array_equal(x[1], r[0][1]) for x in r)
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, blob["P1_S"][:3], blob["P1_U"].dtype, os.path.getsize(path) // 1024, "KiB")


def main():
    outdir = os.path.join(HERE, "..", "tests", "golden")
    if "--dmd" in sys.argv:
        return main_dmd(outdir)
    if "--dtypes" in sys.argv:         # complex128 / float32 fixtures under tests/golden/aux/
        return main_dtypes(outdir)
    os.makedirs(outdir, exist_ok=True)
    if "--rsvd" in sys.argv:           # only the randomized fixtures (leaves the others untouched)
        return main_rsvd(outdir)
    main_rsvd(outdir)
    main_dmd(outdir)
    main_dtypes(outdir)
    for name, m, n, kind, seed, ranks in CASES:
        A = make_input(m, n, kind, seed)
        blob = {"A": A}
        for P in ranks:
            shards = split_rows(A, P)
            r = run_ranks(lambda ref, Ai: ref.svd.tsqr_svd(Ai), shards)
            blob[f"tsqr_svd_P{P}_U"] = np.vstack([x[0] for x in r])
            blob[f"tsqr_svd_P{P}_S"] = r[0][1]
            blob[f"tsqr_svd_P{P}_V"] = r[0][2]
            for x in r[1:]:
                assert np.array_equal(x[1], r[0][1]), "S differs across ranks"
            def pod(ref, Xi):
                U, S, V = ref.POD.run(Xi, remove_mean=True)
                Ur, Sr, Vr = ref.POD.truncate(U, S, V, r=1e-6)
                Xr = ref.POD.reconstruct(Ur, Sr, Vr)
                mean = ref.averaging.temporal_mean(Xi)
                Y = ref.averaging.subtract_mean(Xi, mean)
                rm = ref.stats.RMSE(Y, Xr)
                return U, S, V, Sr.shape[0], Xr, mean, rm
            r = run_ranks(pod, shards)
            blob[f"pod_P{P}_U"] = np.vstack([x[0] for x in r])
            blob[f"pod_P{P}_S"] = r[0][1]
            blob[f"pod_P{P}_V"] = r[0][2]
            blob[f"pod_P{P}_N"] = np.array(r[0][3])
            blob[f"pod_P{P}_Xrec"] = np.vstack([x[4] for x in r])
            blob[f"pod_P{P}_mean"] = np.concatenate([x[5] for x in r])
            blob[f"pod_P{P}_rmse"] = np.array(float(r[0][6]))
        # truncation rule probes on the P=1 spectrum
        ref = _load_rank(_World(1), 0)
        S = blob["tsqr_svd_P1_S"]
        blob["trunc_r"] = np.array([1e-8, 1e-3, 0.5, -0.9, -0.5, -1.0])
        blob["trunc_N"] = np.array([ref.truncation.compute_truncation_residual(S, r) for r in blob["trunc_r"]])
        path = os.path.join(outdir, name + ".npz")
        np.savez_compressed(path, **blob)
        print(name, {k: v.shape for k, v in blob.items() if k.endswith("_S")}, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

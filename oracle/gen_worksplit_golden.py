"""Pins `worksplit` to the reference's own function.  TEST INFRASTRUCTURE (runs in the build container only).

Extracts the source of `worksplit` from the unmodified /root/reference/pyLOM/utils/parall.py (lines 24-48; the module itself
cannot be imported: mpi4py is absent), executes that source text as is, and records its outputs on a grid of
(istart, iend, rank, size) into tests/golden/aux/worksplit_ref.npz.  tests/test_oracle.py checks the oracle's restatement and
the product's `utils.worksplit` against this table (and gen_golden.py's row splits therefore agree with the reference's)."""
import ast, os, sys
import numpy as np

REF = os.environ.get("PYLOM_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
src = open(os.path.join(REF, "pyLOM", "utils", "parall.py")).read()
fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "worksplit")
code = "\n".join(src.splitlines()[fn.lineno - 1:fn.end_lineno])
ns = {"np": np, "MPI_SIZE": 1}
exec(compile(code, "pyLOM/utils/parall.py:worksplit", "exec"), ns)
worksplit = ns["worksplit"]

rows = []
for size in (1, 2, 3, 4, 5, 7, 8, 16):
    for (i0, i1) in ((0, 0), (0, 1), (0, 2), (0, 3), (0, 7), (0, 8), (0, 9), (0, 100), (0, 89351), (0, 8_000_000), (0, 1_000_000_007),
                     (5, 5), (5, 6), (5, 12), (5, 13), (3, 1000), (100, 100 + 192_000_000)):
        for rank in range(size):
            a, b = worksplit(i0, i1, rank, size)
            rows.append((i0, i1, rank, size, int(a), int(b)))
out = os.path.join(HERE, "..", "tests", "golden", "aux", "worksplit_ref.npz")
np.savez_compressed(out, table=np.array(rows, dtype=np.int64))
print(len(rows), "cases ->", os.path.normpath(out))

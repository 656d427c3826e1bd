"""Transcribes outputs the REFERENCE ITSELF printed — the executed cells of its demo notebook
(/root/reference/notebooks/polars_ols_demo.ipynb) — into tests/golden/notebook_outputs.json.

The notebook's data come from `np.random.default_rng(0)` (cell 1: x ~ N(0,1) [2000, 3], eps ~ N(0, 0.1),
y = -sum(x) + eps, group = integers(0, 5), sample_weights = uniform(0, 1)), which is reproducible anywhere, so every
number below is a known answer of the Rust / LAPACK implementation on inputs the tests can regenerate bit for bit
(`notebook_frame()` in tests/test_oracle.py).  Printed precision: 6 decimals.  The script parses the stored cell
outputs (polars' table rendering); it needs /root/reference and is therefore run in the build container only:
    python tests/golden/make_golden_notebook.py
Cells used (0-based index in the .ipynb):
   7  ols(svd, drop) predictions .over(group) / whole frame, wls predictions masked to group 2 — tail(10)
  11  ols(x1, x2, x3, add_intercept=True, mode="coefficients").over("group") — first rows, unnested
  30  ols(solve_method="chol") on an exactly collinear frame -> {null, null, null}
  36  elastic_net(alpha=1e-4, l1_ratio=0.5, positive=True) and ridge(alpha=100, sample_weights) coefficients
  47  rolling_ols(window 252, min_periods 5, alpha 1e-4).over(group), rls(half_life 21, prior mean -1, cov 10)
      .over(group) coefficients, expanding_ols predictions — head(5) / tail(5)
  49  ols(x1, x2, mode="coefficients").over("group") — one row per group
  50  predict(x1, x2) on a fresh 5-feature test frame joined on group — head(5)
"""
import json
import re
from pathlib import Path

NB = Path("/root/reference/notebooks/polars_ols_demo.ipynb")
cells = json.loads(NB.read_text())["cells"]


def output_text(i):
    out = []
    for o in cells[i].get("outputs", []):
        if "text" in o:
            out.append("".join(o["text"]))
        elif "data" in o and "text/plain" in o["data"]:
            out.append("".join(o["data"]["text/plain"]))
    return "\n".join(out)


def table_rows(text, which=0):
    """rows of the `which`-th polars table in `text` as lists of cell strings; wrapped cells are re-joined."""
    tables, cur, in_body = [], None, False
    for line in text.splitlines():
        if line.startswith("╞"):
            cur, in_body = [], True
            continue
        if line.startswith("└"):
            if cur is not None:
                tables.append(cur)
            cur, in_body = None, False
            continue
        if in_body and line.startswith("│"):
            cellsv = [c.strip() for c in re.split("[┆]", line.strip("│"))]
            if cellsv and cellsv[0] == "" and cur:                    # continuation of a wrapped row
                cur[-1] = [a + b for a, b in zip(cur[-1], cellsv)]
            else:
                cur.append(cellsv)
    return tables[which]


def num(s):
    return None if s in ("null", "…") else float(s)


def struct(s):
    return [num(t) for t in s.strip("{}").split(",")]


g = {"source": "azmyrajab/polars_ols notebooks/polars_ols_demo.ipynb (executed outputs stored in the file)",
     "data": "cell 1: rng = np.random.default_rng(0); x = rng.normal(size=(2000, 3)); eps = rng.normal(size=2000, scale=0.1); "
             "y = -x.sum(1) + eps; group = rng.integers(0, 5, size=2000); sample_weights = rng.uniform(0, 1, size=2000)",
     "printed_abs_tol": 1.0e-6}

rows = table_rows(output_text(7))
assert len(rows) == 10
g["cell7_tail10"] = {"x1": [num(r[0]) for r in rows], "predictions_ols_group": [num(r[-3]) for r in rows],
                     "predictions_ols": [num(r[-2]) for r in rows], "predictions_wls_masked": [num(r[-1]) for r in rows]}

rows = table_rows(output_text(11), 1)
g["cell11_head5_unnested"] = {"group": [int(r[0]) for r in rows], "coefficients": [[num(v) for v in r[1:]] for r in rows]}

rows = table_rows(output_text(30))
g["cell30_collinear_chol"] = struct(rows[0][0])

rows = table_rows(output_text(36))
g["cell36"] = {"coef_enet_non_negative": struct(rows[0][0]), "coef_ridge_alpha100_weighted": struct(rows[0][1])}

rows = [r for r in table_rows(output_text(47)) if r[0] != "…"]
assert len(rows) == 10
g["cell47"] = {"rolling_ridge_coef_head5": [struct(r[0]) for r in rows[:5]], "rolling_ridge_coef_tail5": [struct(r[0]) for r in rows[5:]],
               "rls_coef_head5": [struct(r[1]) for r in rows[:5]], "rls_coef_tail5": [struct(r[1]) for r in rows[5:]],
               "expanding_ols_pred_head5": [num(r[2]) for r in rows[:5]], "expanding_ols_pred_tail5": [num(r[2]) for r in rows[5:]]}

rows = table_rows(output_text(49))
g["cell49_by_group"] = {str(int(r[0])): struct(r[1]) for r in rows}

rows = table_rows(output_text(50))
g["cell50_predictions_test_head5"] = [num(r[1]) for r in rows]

out = Path(__file__).parent / "notebook_outputs.json"
out.write_text(json.dumps(g, indent=1) + "\n")
print(f"wrote {out}")

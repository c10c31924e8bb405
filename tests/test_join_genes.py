"""CPU: the string half of transfer_fields (skgenome/combiners.py:58-94 join_strings over bin ranges) -- the native
ordered-unique pass (cnvk_ordered_unique_ranges, host C++) against a direct Python statement of the rule."""
import numpy as np
import pandas as pd

from cnvkit_b200 import params
from cnvkit_b200.segmentation import GeneCodes, _join_genes


def _want(genes, lo, hi, ignore):
    out = []
    for a, b in zip(lo, hi):
        seen = {}
        for g in genes[a:b]:
            if not isinstance(g, str):
                continue
            for part in g.split(","):
                if part not in ignore:
                    seen.setdefault(part, None)
        out.append(",".join(seen) or "-")
    return out


def test_join_genes_matches_the_python_rule():
    rng = np.random.default_rng(3)
    ignore = set(params.IGNORE_GENE_NAMES + params.ANTITARGET_ALIASES)
    names = [f"G{k}" for k in range(40)] + ["Antitarget", "-", "A,B", "B,C", float("nan")]
    for simple in (True, False):
        pool = names[:42] + [float("nan")] if simple else names
        genes = np.array([pool[i] for i in rng.integers(0, len(pool), 3000)], dtype=object)
        genes[100:400] = "G1"                                   # long runs and re-appearances
        genes[900:905] = "Antitarget"
        cuts = np.sort(rng.choice(np.arange(1, 3000), 30, replace=False))
        lo = np.concatenate([[0], cuts, [50]])
        hi = np.concatenate([cuts, [3000], [50]])               # last range is empty
        for col in (genes, pd.Series(genes)):
            gc = GeneCodes(col)
            assert gc.simple == simple
            assert _join_genes(None, lo, hi, gc) == _want(genes, lo, hi, ignore)
    # nested-bin mask: only rows whose end lies beyond the segment start
    genes = np.array(list("ABCDEF"), dtype=object)
    end = np.array([1000, 200, 500, 700, 800, 1200])
    got = _join_genes(genes, np.array([0, 0]), np.array([3, 6]), None, mask_end=end, mask_start=np.array([0, 550]))
    assert got == ["A,B,C", "A,D,E,F"]


def test_unify_levels_restatement():
    """UnifyLevels (haar.py:615-663): addon breakpoints within `window` (inclusive) of a base breakpoint are dropped."""
    from cnvkit_b200.segmentation.haar import unify_levels

    base = np.array([10, 50, 90])
    addon = np.array([5, 9, 11, 12, 49, 51, 70, 89, 91, 92])
    assert unify_levels(base, addon, 1).tolist() == [5, 10, 12, 50, 70, 90, 92]
    assert unify_levels(base, addon, 2).tolist() == [5, 10, 50, 70, 90]
    assert unify_levels(np.array([], dtype=int), addon, 1).tolist() == addon.tolist()
    assert unify_levels(base, np.array([], dtype=int), 1).tolist() == base.tolist()

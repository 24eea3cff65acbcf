"""Static checks of bench.py's multi-rank control flow (no GPU needed).

Regression guard: `_finetune_variant` all-reduces a gradient bucket. Round 1 called it from the rank-0-only block and
a torchrun launch dead-locked against the peers' final barrier (observed once at N=2, 600 s). It is now entered by
EVERY rank, so it must never sit under a rank condition; the remaining rank-0-only extras must stay collective-free.
"""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _calls(node, name):
    return [n for n in ast.walk(node) if isinstance(n, ast.Call) and getattr(n.func, "id", None) == name]


def test_finetune_variant_is_entered_by_every_rank():
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    run_ours = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_ours")
    assert _calls(run_ours, "_finetune_variant"), "bench.py no longer calls _finetune_variant from run_ours?"
    for node in ast.walk(run_ours):
        if isinstance(node, ast.If) and _calls(node, "_finetune_variant"):
            test = ast.get_source_segment(src, node.test)
            assert "rank" not in test and "world == 1" not in test, f"_finetune_variant under a rank condition: {test}"
    # the helpers that only rank 0 runs must not reach a collective themselves
    for fn in ("_time_module_train", "_cudnn_comparator", "_parity_check"):
        node = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == fn)
        for n in ast.walk(node):
            if isinstance(n, ast.Attribute) and n.attr in ("barrier", "all_reduce", "broadcast"):
                raise AssertionError(f"collective `{n.attr}` inside rank-0-only helper {fn} (line {n.lineno})")
        assert not _calls(node, "_barrier") and not _calls(node, "_max_over_ranks"), fn


def test_collectives_in_run_ours_are_not_under_rank_conditions():
    """dist.barrier / all_reduce helpers must be reached by every rank: none may sit under `if rank == 0`."""
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    run_ours = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_ours")
    for node in ast.walk(run_ours):
        if isinstance(node, ast.If) and "rank == 0" in (ast.get_source_segment(src, node.test) or ""):
            for name in ("_barrier", "_max_over_ranks"):
                assert not _calls(node, name), f"{name} under a rank-0 condition"
            for n in ast.walk(node):
                if isinstance(n, ast.Attribute) and n.attr in ("barrier", "all_reduce", "allreduce", "broadcast"):
                    raise AssertionError(f"collective `{n.attr}` under a rank-0 condition (line {n.lineno})")

"""Static checks of bench.py's multi-rank control flow (no GPU needed).

Regression guard: the secondary timings run inside the rank-0-only block; `_finetune_variant` all-reduces a gradient
bucket, so entered by one rank of a torchrun launch it dead-locks against the peers' final barrier (observed once at
N=2, 600 s). They must stay behind a `world == 1` condition.
"""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _calls(node, name):
    return [n for n in ast.walk(node) if isinstance(n, ast.Call) and getattr(n.func, "id", None) == name]


def test_rank0_only_extras_are_single_process_only():
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    run_ours = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "run_ours")
    guarded = []
    for node in ast.walk(run_ours):
        if isinstance(node, ast.If) and _calls(node, "_finetune_variant"):
            guarded.append(ast.get_source_segment(src, node.test))
    assert guarded, "bench.py no longer calls _finetune_variant from run_ours?"
    innermost = min(guarded, key=len)
    assert any("world == 1" in g for g in guarded), f"extras not restricted to one process: {innermost}"


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

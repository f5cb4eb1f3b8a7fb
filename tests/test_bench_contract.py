"""bench.py under torchrun: rank 0 runs extra (side) measurements alone, so nothing that is a collective may sit inside a
rank-0-only block — a barrier there dead-locks every N > 1 run (it did once: the streaming probe's barrier). Checked on the
source, no GPU needed."""
import ast
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COLLECTIVES = {"barrier", "max_over_ranks", "sum_over_ranks", "gather_and_merge", "all_reduce", "all_gather", "gather_object", "broadcast"}


def _mentions_rank0(test: ast.AST) -> bool:
    for n in ast.walk(test):
        if isinstance(n, ast.Compare) and isinstance(n.left, ast.Name) and n.left.id == "rank" and any(isinstance(o, ast.Eq) for o in n.ops):
            if any(isinstance(c, ast.Constant) and c.value == 0 for c in n.comparators):
                return True
    return False


def _calls(node: ast.AST):
    for n in ast.walk(node):
        if isinstance(n, ast.Call):
            f = n.func
            name = f.id if isinstance(f, ast.Name) else f.attr if isinstance(f, ast.Attribute) else None
            yield name, n


def test_no_collective_inside_rank0_only_blocks():
    tree = ast.parse(open(os.path.join(ROOT, "bench.py")).read())
    bad = []
    for node in ast.walk(tree):
        if isinstance(node, ast.If) and _mentions_rank0(node.test):
            for stmt in node.body:
                for name, call in _calls(stmt):
                    if name in COLLECTIVES:
                        bad.append((name, call.lineno))
                    if any(k.arg == "all_ranks" and isinstance(k.value, ast.Constant) and k.value.value for k in call.keywords):
                        bad.append(("all_ranks=True", call.lineno))
    # the merge result is only LOOKED at on rank 0 (merged.size ...): the gather itself is outside
    assert not bad, f"collectives inside rank-0-only blocks of bench.py: {bad}"


def test_streaming_probe_only_synchronises_ranks_when_every_rank_runs_it():
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    fn = next(n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef) and n.name == "run_stream")
    barriers = [c for name, c in _calls(fn) if name == "barrier"]
    assert len(barriers) == 1
    guarded = [n for n in ast.walk(fn) if isinstance(n, ast.If) and isinstance(n.test, ast.Name) and n.test.id == "all_ranks"]
    assert guarded and any(c is b for g in guarded for _, c in _calls(g) for b in barriers)

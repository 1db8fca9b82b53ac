"""Walk a GPU-tier test file on a machine WITHOUT a GPU: every test body runs in dry-run mode
(cupy_b200/_core/_dryrun.py: host logic + codegen + NVRTC compile for sm_100a, nothing launched) with its value
assertions disabled -- device results do not exist here -- while dtype / shape assertions that do not depend on
device data, `pytest.raises` expectations and every host-side exception stay live.

Use: python scripts/dry_gpu_tests.py tests/test_elementwise_family_gpu.py [-k substring]
What it proves: the calls the GPU tier will make are accepted by the host path and every kernel they need
compiles.  With CUPY_B200_CACHE_DIR pointing into the tree it also leaves the cubins behind, so a later run of
the same file on the GPU box (from the same path) starts warm.  TEST INFRASTRUCTURE, never a product path."""
import ast
import inspect
import itertools
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import pytest  # noqa: E402


class _SoftAsserts(ast.NodeTransformer):
    """assert X  ->  try: X  except Exception: pass   (X is still evaluated: it may launch kernels)"""

    def visit_Assert(self, node):
        body = [ast.Expr(node.test)]
        handler = ast.ExceptHandler(type=ast.Name('Exception', ast.Load()), name=None, body=[ast.Pass()])
        return ast.copy_location(ast.fix_missing_locations(ast.Try(body=body, handlers=[handler], orelse=[], finalbody=[])), node)


def _load(path):
    src = open(path).read()
    tree = _SoftAsserts().visit(ast.parse(src, path))
    ast.fix_missing_locations(tree)
    mod = type(sys)('dry_' + os.path.basename(path)[:-3])
    mod.__file__ = path
    exec(compile(tree, path, 'exec'), mod.__dict__)
    return mod


def _cases(fn):
    """Cartesian product of the function's parametrize marks -> list of kwargs."""
    marks = [m for m in getattr(fn, 'pytestmark', []) if m.name == 'parametrize']
    axes = []
    for m in marks:
        names = [n.strip() for n in m.args[0].split(',')]
        vals = []
        for v in m.args[1]:
            v = v.values if hasattr(v, 'values') else v
            vals.append(dict(zip(names, v if len(names) > 1 else (v,))))
        axes.append(vals)
    out = []
    for combo in itertools.product(*axes):
        kw = {}
        for c in combo:
            kw.update(c)
        out.append(kw)
    return out or [{}]


def main():
    path = sys.argv[1]
    sub = sys.argv[sys.argv.index('-k') + 1] if '-k' in sys.argv else ''
    import cupy_b200
    from cupy_b200._core import _dryrun, _ndarray

    def fake_get(self, stream=None, order='C', out=None, blocking=True):
        return np.zeros(self.shape, self.dtype)
    _ndarray.ndarray.get = fake_get
    fake_item = int(os.environ.get('DRY_ITEM', '0'))       # what a device scalar 'reads' as (e.g. a compaction's hit count)
    _ndarray.ndarray.item = lambda self: np.asarray(fake_item).astype(self.dtype).item()
    for name in ('assert_array_equal', 'assert_allclose', 'assert_equal'):
        setattr(np.testing, name, lambda *a, **k: None)

    mod = _load(path)
    ran = failed = 0
    with _dryrun.dry_run() as log:
        for name, fn in sorted(vars(mod).items()):
            if not name.startswith('test_') or not callable(fn) or sub not in name:
                continue
            params = inspect.signature(fn).parameters
            for kw in _cases(fn):
                if 'cp' in params:
                    kw = dict(kw, cp=cupy_b200)
                ran += 1
                try:
                    fn(**kw)
                except pytest.skip.Exception:
                    pass
                except pytest.fail.Exception as e:
                    # e.g. a `pytest.raises` whose exception depends on device DATA (there is none here)
                    print('DATA-DEPENDENT %s: %s' % (name, e))
                except Exception:
                    failed += 1
                    print('FAILED %s %s' % (name, {k: v for k, v in kw.items() if k != 'cp'}))
                    traceback.print_exc()
        kernels = len({d.get('name') for d in log})
    print('%d cases walked, %d failed, %d distinct kernels compiled' % (ran, failed, kernels))
    return 1 if failed else 0


if __name__ == '__main__':
    sys.exit(main())

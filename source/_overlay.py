"""Locating the reference checkout and executing one of its `source/*.py` modules inside an overlay module."""
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)


def reference_root():
    """$TELOGATOR2_REFERENCE, else a checkout next to this repository (telogator2/, reference/), else /root/reference."""
    cands = [os.environ.get('TELOGATOR2_REFERENCE'), os.path.join(os.path.dirname(_ROOT), 'telogator2'),
             os.path.join(os.path.dirname(_ROOT), 'reference'), '/root/reference']
    for c in cands:
        if c and os.path.isfile(os.path.join(c, 'source', 'tg_align.py')) and os.path.abspath(os.path.join(c, 'source')) != _HERE:
            return os.path.abspath(c)
    raise ImportError('telogator2_b200 `source` overlay: reference checkout not found; set TELOGATOR2_REFERENCE to the directory '
                      'that holds telogator2.py and source/')


def exec_reference(module_name, namespace):
    """Run <reference>/source/<module>.py in `namespace` (the overlay module's globals): afterwards the overlay module IS the
    reference module, and its `from source.X import ...` lines resolved against the overlay package."""
    short = module_name.rsplit('.', 1)[-1]
    path = os.path.join(reference_root(), 'source', short + '.py')
    with open(path, 'r') as f:
        code = compile(f.read(), path, 'exec')
    namespace['__reference_file__'] = path
    exec(code, namespace)

"""nphm_b200 - B200-native engine for the hot path of NPHM (Neural Parametric Head Models).

The sub-packages mirror the reference's module layout (``NPHM.models.*``, ``NPHM.utils.*``) so that the
reference's ``scripts/fitting`` and ``scripts/training`` run unchanged on top of this engine after
:func:`install_as_nphm`.
"""
import sys as _sys

__version__ = '0.1.0'


def install_as_nphm(force: bool = False):
    """Register this package's mirrors under the reference's import names
    (``NPHM.models.EnsembledDeepSDF`` ...), so ``from NPHM.models.deepSDF import DeepSDF`` resolves here."""
    import importlib
    import types
    if 'NPHM' in _sys.modules and not force and not getattr(_sys.modules['NPHM'], '_nphm_b200_alias', False):
        raise RuntimeError('a different NPHM package is already imported; pass force=True to shadow it')
    root = types.ModuleType('NPHM')
    root._nphm_b200_alias = True
    root.__path__ = []
    _sys.modules['NPHM'] = root
    for sub in ('models', 'utils'):
        pkg = importlib.import_module('nphm_b200.' + sub)
        _sys.modules['NPHM.' + sub] = pkg
        setattr(root, sub, pkg)
    for name in ('models.EnsembledDeepSDF', 'models.deepSDF', 'models.reconstruction', 'models.fitting',
                 'models.iterative_root_finding', 'models.diff_operators', 'utils.reconstruction'):
        try:
            mod = importlib.import_module('nphm_b200.' + name)
        except ImportError:
            continue
        _sys.modules['NPHM.' + name] = mod
    return root

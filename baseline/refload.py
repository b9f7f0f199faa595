"""Imports the unmodified reference modules (models, loss_fns, utils, config) from baseline/_ref -- or straight from
/root/reference when it is mounted -- with the two shims the reference needs in this image: a stub for
`matplotlib.colors` (utils.py:9; only used by utils.augmentation, which main.py:56 has commented out) and
nothing else.  Returns None when the reference is not available."""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
_cached = None


def reference_dir():
    for d in (os.path.join(HERE, '_ref'), os.environ.get('CAPS_REFERENCE_DIR', '/root/reference')):
        if d and os.path.exists(os.path.join(d, 'models.py')):
            return d
    return None


def load():
    """-> namespace with .models .loss_fns .utils .config .dir, or None."""
    global _cached
    if _cached is not None:
        return _cached
    d = reference_dir()
    if d is None:
        return None
    if 'matplotlib' not in sys.modules:
        try:
            import matplotlib.colors  # noqa: F401
        except Exception:
            mpl = types.ModuleType('matplotlib')
            col = types.ModuleType('matplotlib.colors')
            col.rgb_to_hsv = lambda a: a
            col.hsv_to_rgb = lambda a: a
            mpl.colors = col
            sys.modules['matplotlib'] = mpl
            sys.modules['matplotlib.colors'] = col
    if d not in sys.path:
        sys.path.insert(0, d)
    ns = types.SimpleNamespace(dir=d)
    for name in ('config', 'utils', 'loss_fns', 'models'):
        mod = importlib.import_module(name)
        if os.path.dirname(os.path.abspath(mod.__file__)) != os.path.abspath(d):
            raise RuntimeError('module %r resolved to %s, not to the reference in %s' % (name, mod.__file__, d))
        setattr(ns, name, mod)
    _cached = ns
    return ns


def make_params(ref, model, device, recon):
    """What main.load_params builds (main.py:227-241), minus the tensorboard writer."""
    p = ref.utils.Params(os.path.join(ref.dir, 'experiments', model, 'params.json'))
    p.device = device
    p.seed = 0
    p.model = model
    p.recon = recon
    p.recon_coef = 5e-4          # main.py:33 default
    p.eval_every = 1
    p.train_frac = 1.0
    return p

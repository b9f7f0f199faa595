"""Brings the UNMODIFIED reference next to the repo so that it can run on the GPU box (where /root/reference does
not exist).

The reference is a flat directory of Python scripts without setup.py / pyproject.toml, so `pip install --target
baseline/_ref /root/reference` has nothing to build; this recipe is its equivalent: it copies the handful of files
the capsule / darkcapsule paths import (byte for byte, verified by sha256) into baseline/_ref/, which is listed in
.gitignore (never enters history) but not in .gpurunignore (travels to the GPU box like a built .so does).
Called by __graft_entry__.build(); a no-op when /root/reference is not mounted (the GPU box uses what travelled)."""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
SRC = os.environ.get('CAPS_REFERENCE_DIR', '/root/reference')
FILES = ['models.py', 'loss_fns.py', 'utils.py', 'config.py',
         'experiments/capsule/params.json', 'experiments/darkcapsule/params.json']


def _sha(path):
    return hashlib.sha256(open(path, 'rb').read()).hexdigest()


def install(force=False):
    """Returns the install directory, or None if the reference is neither mounted nor already installed."""
    manifest = os.path.join(DST, 'MANIFEST.json')
    if not os.path.isdir(SRC):
        return DST if os.path.exists(manifest) else None
    want = {f: _sha(os.path.join(SRC, f)) for f in FILES}
    if not force and os.path.exists(manifest):
        try:
            if json.load(open(manifest)).get('sha256') == want and all(os.path.exists(os.path.join(DST, f)) for f in FILES):
                return DST
        except Exception:
            pass
    for f in FILES:
        dst = os.path.join(DST, f)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, f), dst)
        assert _sha(dst) == want[f]
    json.dump({'source': SRC, 'sha256': want}, open(manifest, 'w'), indent=1)
    return DST


if __name__ == '__main__':
    print(install(force=True))

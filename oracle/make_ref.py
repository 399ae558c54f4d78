"""Recipe: stage the reference's OWN importable modules for the score/SDE/sampler path into oracle/_ref/.

TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py).  The reference (moonbow721/DPoser) is pure Python:
there is nothing to compile.  ``build()`` runs this in the build container, where /root/reference exists; the
files land in the git-ignored ``oracle/_ref/`` (never committed -- reference sources stay out of the history) and
travel to the GPU box with the snapshot, where ``bench.py --impl reference`` and the ``cpu_baseline`` leg import
them UNMODIFIED (``sys.path`` gets ``oracle/_ref``) and time ``lib.algorithms.advanced.sampling.get_pc_sampler``
itself on the host cores.  Third-party ``smplx`` (the LBS arithmetic) is absent from the image, so the LBS part of
the baseline stays the oracle port (oracle/lbs_ref.py) -- bench.py says so in ``cpu_baseline.sample``.
"""
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
DST = os.path.join(HERE, '_ref')
FILES = ['lib/algorithms/advanced/model.py', 'lib/algorithms/advanced/sde_lib.py', 'lib/algorithms/advanced/utils.py',
         'lib/algorithms/advanced/sampling.py', 'lib/algorithms/advanced/likelihood.py', 'lib/algorithms/ema.py']


def available():
    return all(os.path.exists(os.path.join(DST, f)) for f in FILES)


def stage(force=False):
    """Copy FILES from /root/reference into oracle/_ref/ (byte-identical).  Returns True if oracle/_ref is usable."""
    if not os.path.isdir(REF):
        return available()
    for f in FILES:
        src, dst = os.path.join(REF, f), os.path.join(DST, f)
        if force or not os.path.exists(dst) or os.path.getmtime(dst) < os.path.getmtime(src):
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
    return available()


def import_reference():
    """(model, sde_lib, utils, sampling) modules of the staged reference, or None when oracle/_ref is absent."""
    import sys
    if not available():
        return None
    if DST not in sys.path:
        sys.path.insert(0, DST)
    from lib.algorithms.advanced import model, sampling, sde_lib, utils
    return model, sde_lib, utils, sampling


if __name__ == '__main__':
    print('oracle/_ref staged:', stage(force=True))

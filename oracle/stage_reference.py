"""Stages the UNMODIFIED reference under oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).

    python oracle/stage_reference.py        (container only: needs /root/reference)

What it does - the base contract's one offline install, pointed at oracle/_ref instead of baseline/_ref:
    python -m pip install --no-index --no-build-isolation --no-deps --target oracle/_ref <copy of /root/reference>
(from a copy under /tmp because /root/reference is read-only and setuptools writes build files beside setup.py),
then gunzips example_dataset/*.csv.gz into oracle/_ref/example_dataset/ (the data `PROJECT_DATA_PATH` points at,
settings.py:8-18).  Nothing of the reference is committed: oracle/_ref/ is listed in .gitignore.

Who uses it (test infrastructure only, like everything under oracle/): `oracle/ref_import.py` (the pinning tests and
the drop-in proof of `oracle/dropin.py`), `bench.py --impl reference` and bench.py's cpu_baseline (the reference's own
numba kernels timed on the GPU box's host cores).  The product package never imports it.
"""
import gzip
import os
import shutil
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = '/root/reference'
TARGET = os.path.join(HERE, '_ref')


def staged():
    return os.path.isfile(os.path.join(TARGET, 'doppelspeller', 'match_maker.py')) and \
        os.path.isfile(os.path.join(TARGET, 'example_dataset', 'example_truth.csv'))


def stage(force=False):
    if staged() and not force:
        return TARGET
    if not os.path.isdir(os.path.join(REFERENCE_ROOT, 'doppelspeller')):
        raise RuntimeError(f'{REFERENCE_ROOT} is not present: the reference can only be staged in the build container')
    shutil.rmtree(TARGET, ignore_errors=True)
    os.makedirs(TARGET)
    work = tempfile.mkdtemp(prefix='ds_ref_src_')
    try:
        source = os.path.join(work, 'reference')
        shutil.copytree(REFERENCE_ROOT, source, ignore=shutil.ignore_patterns('.git', 'description.jpg'))
        cmd = [sys.executable, '-m', 'pip', 'install', '--no-index', '--no-build-isolation', '--no-deps', '--quiet',
               '--find-links', '/opt/wheelhouse', '--target', TARGET, source]
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=dict(os.environ, PROJECT_DATA_PATH=work))
        if proc.returncode != 0:
            raise RuntimeError('pip install of the reference failed:\n' + proc.stdout[-3000:])
    finally:
        shutil.rmtree(work, ignore_errors=True)
    data = os.path.join(TARGET, 'example_dataset')
    os.makedirs(data, exist_ok=True)
    for name in sorted(os.listdir(os.path.join(REFERENCE_ROOT, 'example_dataset'))):
        if name.endswith('.csv.gz'):
            with gzip.open(os.path.join(REFERENCE_ROOT, 'example_dataset', name), 'rb') as fin, \
                    open(os.path.join(data, name[:-3]), 'wb') as fout:
                shutil.copyfileobj(fin, fout)
    if not staged():
        raise RuntimeError('staging finished but oracle/_ref is incomplete')
    return TARGET


if __name__ == '__main__':
    print(stage(force='--force' in sys.argv))

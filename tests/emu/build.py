"""TEST INFRASTRUCTURE ONLY: builds tests/emu/_build/libds_emu[_asan].so - the kernel sources of
doppelspeller_b200/csrc, translated by translate.py and compiled for the HOST against cuda_emu.h (see its header).

    python tests/emu/build.py            # plain -O2 build
    python tests/emu/build.py --asan     # -fsanitize=address,undefined (run python with LD_PRELOAD=libasan.so)
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, 'doppelspeller_b200', 'csrc')
BUILD = os.path.join(HERE, '_build')
SOURCES = ('ds_topn.cu', 'ds_pairs.cu', 'ds_encode.cu', 'ds_gbdt.cu')


def _load_translate():
    import importlib.util
    spec = importlib.util.spec_from_file_location('ds_emu_translate', os.path.join(HERE, 'translate.py'))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


translate = _load_translate()


def library_path(asan, tsan=False):
    return os.path.join(BUILD, 'libds_emu_tsan.so' if tsan else 'libds_emu_asan.so' if asan else 'libds_emu.so')


def asan_runtime(name='libasan.so'):
    """Path of libasan.so / libtsan.so (must be LD_PRELOADed into python), or None when the toolchain has none."""
    try:
        path = subprocess.run(['gcc', f'-print-file-name={name}'], stdout=subprocess.PIPE, text=True, check=True).stdout.strip()
    except (OSError, subprocess.CalledProcessError):
        return None
    return path if os.path.isabs(path) and os.path.exists(path) else None


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(asan=False, verbose=False, tsan=False):
    out = library_path(asan, tsan)
    inputs = [os.path.join(CSRC, s) for s in SOURCES] + [os.path.join(CSRC, 'ds_common.cuh'),
                                                         os.path.join(ROOT, 'include', 'doppelspeller_b200.h')]
    inputs += [os.path.join(HERE, f) for f in ('cuda_emu.h', 'cuda_emu.cpp', 'translate.py', 'build.py')]
    inputs += [os.path.join(HERE, 'include', 'cub', 'cub_emu.cuh')]
    if os.path.exists(out) and os.path.getmtime(out) >= _newest(inputs):
        return out
    tag = 'tsan' if tsan else 'asan' if asan else 'plain'
    work = os.path.join(BUILD, tag)
    os.makedirs(work, exist_ok=True)
    # -ffp-contract=off = nvcc --fmad=false; -frounding-math keeps the directed-rounding intrinsics honest
    flags = ['-std=c++17', '-fPIC', '-g1', '-ffp-contract=off', '-frounding-math', '-fno-strict-aliasing', '-w',
             '-I', os.path.join(HERE, 'include'), '-I', CSRC, '-include', os.path.join(HERE, 'cuda_emu.h')]
    engine_flags = list(flags)
    if tsan:
        flags += ['-O1', '-fno-omit-frame-pointer', '-fsanitize=thread']
        engine_flags += ['-O2', '-DDS_EMU_TSAN']          # the engine itself is NOT instrumented: it calls the TSan fiber API
    elif asan:
        flags += ['-O1', '-fno-omit-frame-pointer', '-fsanitize=address,undefined', '-fno-sanitize-recover=undefined',
                  '-fno-sanitize=vptr']
        engine_flags = flags
    else:
        flags += ['-O2']
        engine_flags = flags
    jobs = []
    for name in SOURCES:
        src = os.path.join(CSRC, name)
        translated = os.path.join(work, name.replace('.cu', '.emu.cpp'))
        with open(src) as f:
            text, launches, dynamic = translate.translate(f.read(), src)
        with open(translated, 'w') as f:
            f.write(text)
        if verbose:
            print(f'{name}: {launches} launches, {dynamic} dynamic shared-memory declarations')
        jobs.append((translated, os.path.join(work, name.replace('.cu', '.o'))))
    jobs.append((os.path.join(HERE, 'cuda_emu.cpp'), os.path.join(work, 'cuda_emu.o')))

    def compile_one(job):
        src, obj = job
        proc = subprocess.run(['g++'] + (engine_flags if src.endswith('cuda_emu.cpp') else flags) + ['-c', src, '-o', obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError(f'g++ failed on {src}:\n{proc.stdout[-6000:]}')
    with ThreadPoolExecutor(max_workers=len(jobs)) as pool:
        list(pool.map(compile_one, jobs))
    link = ['g++', '-shared', '-o', out] + [obj for _, obj in jobs]
    if tsan:
        link += ['-fsanitize=thread']
    elif asan:
        link += ['-fsanitize=address,undefined']
    proc = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if proc.returncode != 0:
        raise RuntimeError(f'link failed:\n{proc.stdout[-4000:]}')
    return out


if __name__ == '__main__':
    print(build(asan='--asan' in sys.argv, tsan='--tsan' in sys.argv, verbose=True))

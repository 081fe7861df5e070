"""CPU: the UNMODIFIED kernel sources of doppelspeller_b200/csrc, compiled for the host and executed thread by thread
(tests/emu/cuda_emu.h: every CUDA thread a fiber, __syncthreads / __syncwarp / shuffles / votes as rendezvous points,
device allocations as exact-size heap blocks filled with garbage), so that

  * the `gpu` parity tests - golden vectors minted from the reference, the CPU oracle, the fuzzers, the sharded phases,
    the shared-threshold exchange - also run HERE, where there is no GPU, against the very code the B200 executes;
  * AddressSanitizer and UBSan check every shared- and global-memory access of the kernels (compute-sanitizer is closed
    on the GPU pool): out-of-bounds smem / global indexes, misaligned vector loads, reads of freed
    workspaces, launches that ask for more shared memory than was opted into, collectives whose lanes disagree,
    barriers that cannot complete, shuffles that read lanes outside the collective, leaked device blocks.

The emulated library is test infrastructure: built into tests/emu/_build/ (git-ignored), loaded only by the
subprocesses below through DOPPELSPELLER_B200_LIB.  The product has no CPU path (tests/test_library_abi.py).
"""
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, 'tests', 'emu')


def _load(name):
    """tests/emu/<name>.py under a private module name (`build` / `translate` are too generic for sys.path)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location(f'ds_emu_{name}', os.path.join(EMU, f'{name}.py'))
    module = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(module)
    return module


emu_build = _load('build')

SANITIZER_REPORT = re.compile(r'ERROR: AddressSanitizer|runtime error:|ds_emu: FATAL|ds_emu: deadlock|ds_emu: invalid launch')
SANITIZER_ENV = {'ASAN_OPTIONS': 'detect_leaks=0:detect_stack_use_after_return=0:abort_on_error=0',
                 'UBSAN_OPTIONS': 'print_stacktrace=1:halt_on_error=1'}
# DS_EMU_FULL=1: every emulated test under the sanitizers (about five minutes); the default keeps the CPU suite short:
# the whole set on the plain build, the edge cases / fuzzers / device-pointer paths under the sanitizers
FULL = os.environ.get('DS_EMU_FULL') == '1'
EMULATED_FILES = ['tests/test_gpu_parity.py', 'tests/test_gpu_dataframe_api.py', 'tests/emu/device_paths.py']
if FULL:      # + the reference's own Prediction with the three imports swapped, on the emulated kernels (needs oracle/_ref)
    EMULATED_FILES.insert(2, 'tests/test_gpu_dropin.py')
SANITIZED_SUBSET = ('edge or randomised or negative_weight or long_title or large_vocabulary or buffer_overflow or sums_match '
                    'or indel_ratio_matches_reference or construct_features or levenshtein or idf_word or transform_titles '
                    'or single_title or prematch or trigram_encoder or title_features or thresholds_shared or pair_kernels '
                    'or absent_lane or 3000-300-1-3 or 700-100-512-5 or 300-5 or 3-12 or 0-1')

pytestmark = pytest.mark.skipif(sys.platform != 'linux' or os.uname().machine != 'x86_64',
                                reason='the fiber switch of tests/emu/cuda_emu.cpp is x86-64 System V assembly')


def _have_compiler():
    try:
        return subprocess.run(['g++', '--version'], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL).returncode == 0
    except OSError:
        return False


def _workers():
    try:
        import xdist  # noqa: F401
    except ImportError:
        return []
    return ['-n', str(max(1, min(4, (os.cpu_count() or 2) // 2)))]


def _run_emulated(library, selection, sanitized):
    """The emulated `gpu` tests in a subprocess (pytest-xdist workers when available: every worker is its own process with
    its own emulated device) -> (tests passed, [launches, CTAs, threads, collectives, absent-lane reads, live blocks] summed
    over the processes).  Sanitizer reports go to files and fail the run."""
    work = os.path.join(EMU, '_build', 'run_asan' if sanitized else 'run_plain')
    os.makedirs(work, exist_ok=True)
    for stale in os.listdir(work):
        os.remove(os.path.join(work, stale))
    env = dict(os.environ, DOPPELSPELLER_B200_LIB=library, DS_EMU_STATS=os.path.join(work, 'stats'))
    if sanitized:
        env.update(ASAN_OPTIONS=SANITIZER_ENV['ASAN_OPTIONS'] + ':log_path=' + os.path.join(work, 'asan'),
                   UBSAN_OPTIONS=SANITIZER_ENV['UBSAN_OPTIONS'] + ':log_path=' + os.path.join(work, 'ubsan'), LD_PRELOAD=emu_build.asan_runtime())
    # -s: whatever a dying process still prints must not be lost inside pytest's capture
    cmd = [sys.executable, '-m', 'pytest', *EMULATED_FILES, '-m', 'gpu', '-p', 'tests.emu.plugin', '-q', '-x', '-s', '-p', 'no:cacheprovider']
    cmd += _workers()
    if selection:
        cmd += ['-k', selection]
    proc = subprocess.run(cmd, cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=3000)
    reports = ''
    for name in sorted(os.listdir(work)):
        if name.startswith(('asan', 'ubsan')):
            with open(os.path.join(work, name)) as f:
                reports += f.read()[:4000]
    tail = proc.stdout[-6000:] + reports
    assert proc.returncode == 0, tail
    assert not reports and not SANITIZER_REPORT.search(proc.stdout), tail
    summary = re.search(r'(\d+) passed', proc.stdout)
    assert summary and not re.search(r'\d+ (failed|error)', proc.stdout), tail
    totals = [0] * 6
    with open(os.path.join(work, 'stats')) as f:
        lines = re.findall(r'ds_emu: (\d+) launches, (\d+) CTAs, (\d+) threads, (\d+) warp collectives, \d+ fiber switches, '
                           r'(\d+) shuffle reads of absent lanes, (\d+) live device blocks', f.read())
    assert lines, tail
    for line in lines:
        totals = [a + int(b) for a, b in zip(totals, line)]
    return int(summary.group(1)), totals


@pytest.mark.skipif(not _have_compiler(), reason='g++ not available')
def test_emulator_reports_planted_faults():
    """The harness is only evidence if it notices errors: kernels with planted faults (tests/emu/selftest.cu) - an index one
    past the dynamic / a static shared array, one past a global buffer, a misaligned 128-bit load, a freed buffer, lanes
    meeting in different collectives, a barrier that cannot complete, > 48 KB of shared memory without the opt-in, an
    empty grid - must each be reported, and the fault-free kernel must pass."""
    if emu_build.asan_runtime() is None:
        pytest.skip('no AddressSanitizer runtime in this toolchain')
    translate = emu_build.translate
    work = os.path.join(EMU, '_build')
    os.makedirs(work, exist_ok=True)
    source = os.path.join(EMU, 'selftest.cu')
    with open(source) as f:
        text, launches, dynamic = translate.translate(f.read(), source)
    assert launches == 13 and dynamic == 2
    translated = os.path.join(work, 'selftest.emu.cpp')
    with open(translated, 'w') as f:
        f.write(text)
    binary = os.path.join(work, 'selftest')
    subprocess.run(['g++', '-std=c++17', '-g1', '-O1', '-fno-omit-frame-pointer', '-fsanitize=address,undefined',
                    '-fno-sanitize-recover=undefined', '-ffp-contract=off', '-frounding-math', '-w', '-I', os.path.join(EMU, 'include'),
                    '-include', os.path.join(EMU, 'cuda_emu.h'), translated, os.path.join(EMU, 'cuda_emu.cpp'), '-o', binary], check=True)
    env = dict(os.environ, **SANITIZER_ENV)
    expected = {
        'clean': (0, r'selftest clean: ok'),
        'dynamic-smem-overflow': (None, r'AddressSanitizer: heap-buffer-overflow'),
        'static-smem-overflow': (None, r"index 64 out of bounds for type 'int \[64\]'"),
        'global-overflow': (None, r'AddressSanitizer: heap-buffer-overflow'),
        'misaligned-vector-load': (None, r'misaligned address .* requires 16 byte alignment'),
        'use-after-free': (None, r'AddressSanitizer: heap-use-after-free'),
        'mismatched-collectives': (None, r'lanes 0 and 16 meet in different collectives'),
        'barrier-deadlock': (None, r'ds_emu: deadlock in kernel k_barrier_deadlock'),
        'smem-without-opt-in': (5, r'invalid launch configuration of k_clean'),
        'empty-grid': (5, r'invalid launch configuration of k_touch'),
    }
    for case, (code, pattern) in expected.items():
        proc = subprocess.run([binary, case], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=120)
        assert re.search(pattern, proc.stdout), (case, proc.stdout[-2000:])
        assert (proc.returncode == code) if code is not None else (proc.returncode != 0), (case, proc.returncode)
    # racecheck: the kernels instrumented by ThreadSanitizer, every CUDA thread a TSan fiber (the engine is not instrumented)
    if emu_build.asan_runtime('libtsan.so') is None:
        return
    common = ['-std=c++17', '-g1', '-ffp-contract=off', '-frounding-math', '-w', '-I', os.path.join(EMU, 'include'), '-include',
              os.path.join(EMU, 'cuda_emu.h')]
    kernels, engine, racecheck = (os.path.join(work, name) for name in ('selftest_tsan.o', 'cuda_emu_tsan.o', 'selftest_tsan'))
    subprocess.run(['g++', *common, '-O1', '-fsanitize=thread', '-c', translated, '-o', kernels], check=True)
    subprocess.run(['g++', *common, '-O2', '-DDS_EMU_TSAN', '-c', os.path.join(EMU, 'cuda_emu.cpp'), '-o', engine], check=True)
    subprocess.run(['g++', '-fsanitize=thread', kernels, engine, '-o', racecheck], check=True)
    races = {'clean': 0, 'ordered-by-syncwarp': 0, 'race-missing-syncthreads': 1, 'race-missing-syncwarp': 1}
    for case, n_races in races.items():
        proc = subprocess.run([racecheck, case], env=dict(os.environ, TSAN_OPTIONS='halt_on_error=0'), stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True, timeout=120)
        assert proc.stdout.count('WARNING: ThreadSanitizer: data race') == n_races, (case, proc.stdout[-3000:])
        if n_races:
            assert re.search(rf'k_{case.replace("-", "_")}\(int\*\) .*selftest.cu:\d+', proc.stdout), (case, proc.stdout[-3000:])


@pytest.mark.skipif(not _have_compiler(), reason='g++ not available')
def test_kernel_parity_on_the_emulated_device():
    """Every `gpu` parity test that works on host buffers, plus the device-pointer paths of tests/emu/device_paths.py,
    against the kernels executed on the host: bit-exact with the golden vectors and the oracle, no shuffle reads a lane
    outside its collective, every workspace is released."""
    library = emu_build.build(asan=False)
    passed, (launches, ctas, threads, collectives, absent, live) = _run_emulated(library, None, sanitized=False)
    assert passed >= 40
    assert launches > 2000 and threads > 50_000_000 and collectives > 5_000_000
    assert absent == 0 and live == 0


@pytest.mark.skipif(not _have_compiler(), reason='g++ not available')
def test_kernels_under_address_and_undefined_behaviour_sanitizers():
    """The same sources built with -fsanitize=address,undefined: the edge cases, the shape fuzzers and the device-pointer
    paths (DS_EMU_FULL=1: everything) must pass without a single sanitizer report."""
    if emu_build.asan_runtime() is None:
        pytest.skip('no AddressSanitizer runtime in this toolchain')
    library = emu_build.build(asan=True)
    passed, (launches, ctas, threads, collectives, absent, live) = _run_emulated(library, None if FULL else SANITIZED_SUBSET, sanitized=True)
    assert passed >= (40 if FULL else 25)
    assert absent == 0 and live == 0


def _statement_line(path, statement, after=None):
    """`file:line` of the one line holding `statement` (the first one after the line holding `after`, when given)"""
    with open(path) as f:
        text = f.readlines()
    start = 0
    if after is not None:
        start = next(i for i, line in enumerate(text) if after in line)
    lines = [i + 1 for i, line in enumerate(text) if i >= start and statement in line]
    assert lines and (after is not None or len(lines) == 1), (path, statement, lines)
    return f'{os.path.basename(path)}:{lines[0]}'


@pytest.mark.skipif(not FULL, reason='about ten minutes: set DS_EMU_FULL=1 (outcome recorded in profiles/r2_emulation.md)')
def test_kernels_under_thread_sanitizer():
    """Racecheck of every emulated test in the STRICT model (only __syncthreads / __syncwarp order memory; votes and shuffles
    do not): no hazard on shared memory anywhere; the only reports are formally unordered global accesses, each benign -
    lanes reading a per-query word that lane 0 of the same warp rewrites on its way out (every lane leaves without side
    effects whichever value it sees: `state`, `cand_count`, `theta`), and k_trigrams' warps all storing the same 1 into
    present[code]."""
    if emu_build.asan_runtime('libtsan.so') is None:
        pytest.skip('no ThreadSanitizer runtime in this toolchain')
    library = emu_build.build(tsan=True)
    log = os.path.join(EMU, '_build', 'tsan_report')
    for stale in [f for f in os.listdir(os.path.dirname(log)) if f.startswith('tsan_report')]:
        os.remove(os.path.join(os.path.dirname(log), stale))
    env = dict(os.environ, DOPPELSPELLER_B200_LIB=library, DS_EMU_STATS='1', LD_PRELOAD=emu_build.asan_runtime('libtsan.so'),
               TSAN_OPTIONS=f'halt_on_error=0:report_signal_unsafe=0:history_size=2:exitcode=0:log_path={log}')
    files = [f for f in EMULATED_FILES if 'dropin' not in f]     # (the reference's numba threads under TSan take an hour)
    proc = subprocess.run([sys.executable, '-m', 'pytest', *files, '-m', 'gpu', '-p', 'tests.emu.plugin', '-q', '-x', '-s', '-p',
                           'no:cacheprovider', *_workers()], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=6000)
    assert proc.returncode == 0 and re.search(r'(\d+) passed', proc.stdout), proc.stdout[-6000:]
    topn, encode = (os.path.join(ROOT, 'doppelspeller_b200', 'csrc', name) for name in ('ds_topn.cu', 'ds_encode.cu'))
    overflow = 'if (cnt > p.cap) {'
    benign = {
        # k_select: words every lane reads on entry and lane 0 rewrites on its way out
        _statement_line(topn, 'if (p.state[b] & STATE_OVERFLOW) return;'), _statement_line(topn, 'p.state[b] |= STATE_OVERFLOW;'),
        _statement_line(topn, 'double theta = by_row ? p.threshold[q] : p.theta[b];'), _statement_line(topn, 'p.theta[b] = ext;'),
        _statement_line(topn, 'int cnt = p.cand_count[b];'), _statement_line(topn, 'p.cand_count[b] = 0;', after=overflow),
        # k_trigrams: every warp stores the same 1
        _statement_line(encode, 'present[x] = 1;'),
    }
    unexpected = []
    for name in os.listdir(os.path.dirname(log)):
        if not name.startswith('tsan_report'):
            continue
        with open(os.path.join(os.path.dirname(log), name)) as f:
            for report in f.read().split('=================='):
                if 'WARNING: ThreadSanitizer' not in report:
                    continue
                sites = frozenset(re.findall(r'#0 [^\n]*?/csrc/(ds_\w+\.cu:\d+)', report))
                if sites and not sites <= benign:          # reports without a kernel frame are the oracle's OpenMP threads
                    unexpected.append(report[:3000])
    assert not unexpected, unexpected[0]

"""Per CUDA source line totals (instructions executed, stall samples, shared-memory wavefronts) of one captured launch
from an `ncu --set full --import-source on` report:  python profiles/source_hot.py report.ncu-rep [launch index] [top n]"""
import csv
import subprocess
import sys


def main():
    report, which, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0, int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(['ncu', '-i', report, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    # one section per (launch, file); a launch starts with the section of the kernel's own .cu file
    launches, current, path = [], None, ''
    for r in rows:
        if r and r[0] == 'File Path':
            path = r[1]
            if path.endswith('.cu'):
                launches.append([None, []])
        elif r and r[0] == 'Line No':
            launches[-1][0] = r
        elif r and r[0] not in ('Function Name',) and r[0] != '' and launches:
            launches[-1][1].append([r[0] if path.endswith('.cu') else path.rsplit('/', 1)[-1] + ':' + r[0]] + r[1:])
    header, lines = launches[which]
    col = {name: i for i, name in enumerate(header)}
    def val(r, name):
        try:
            return float(r[col[name]])
        except (ValueError, KeyError, IndexError):
            return 0.0
    total_inst = sum(val(r, 'Instructions Executed') for r in lines)
    total_samples = sum(val(r, '# Samples') for r in lines)
    total_wave = sum(val(r, 'L1 Wavefronts Shared') for r in lines)
    print(f'launch {which}: {total_inst:.0f} warp instructions, {total_samples:.0f} samples, {total_wave:.0f} shared wavefronts')
    print('line | inst share | sample share | smem wavefront share | source')
    for r in sorted(lines, key=lambda r: -val(r, '# Samples'))[:top]:
        print(f'{r[0]:>5} | {val(r, "Instructions Executed") / max(1, total_inst):6.3f} | {val(r, "# Samples") / max(1, total_samples):6.3f} | '
              f'{val(r, "L1 Wavefronts Shared") / max(1, total_wave):6.3f} | {r[1].strip()[:110]}')


if __name__ == '__main__':
    main()

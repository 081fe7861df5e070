"""TEST INFRASTRUCTURE ONLY: rewrites the two pieces of CUDA syntax g++ cannot parse, so that the kernel sources of
doppelspeller_b200/csrc compile unchanged otherwise against tests/emu/cuda_emu.h:

    kernel<<<grid, block, smem, stream>>>(args);      ->  ::ds_emu::launch(&kernel, "kernel", dim3(grid), dim3(block), smem, stream,
                                                                           [&]() { kernel(args); });
    extern __shared__ __align__(16) unsigned char s[]; ->  unsigned char *s = ::ds_emu::dynamic_smem();

Line numbers are preserved (a `#line` directive names the original file), so sanitizer reports point into the real
sources."""
import re
import sys

LAUNCH = re.compile(r'(?<![\w:])([A-Za-z_]\w*(?:\s*<[^<>;(){}]*>)?)\s*<<<')
DYNAMIC_SMEM = re.compile(r'extern\s+__shared__\s+(?:__align__\(\s*\d+\s*\)\s+)?unsigned\s+char\s+(\w+)\s*\[\s*\]\s*;')


def _split_top_level(text):
    parts, depth, start = [], 0, 0
    for i, ch in enumerate(text):
        if ch in '([{':
            depth += 1
        elif ch in ')]}':
            depth -= 1
        elif ch == ',' and depth == 0:
            parts.append(text[start:i])
            start = i + 1
    parts.append(text[start:])
    return [p.strip() for p in parts]


def _matching_paren(text, open_at):
    depth = 0
    for i in range(open_at, len(text)):
        if text[i] == '(':
            depth += 1
        elif text[i] == ')':
            depth -= 1
            if depth == 0:
                return i
    raise ValueError('unbalanced parentheses after a kernel launch')


def translate(source, path):
    out, at, launches = [], 0, 0
    while True:
        m = LAUNCH.search(source, at)
        if m is None:
            break
        kernel = re.sub(r'\s+', '', m.group(1))
        close = source.index('>>>', m.end())
        config = _split_top_level(source[m.end():close])
        if not 2 <= len(config) <= 4:
            raise ValueError(f'{path}: launch of {kernel} with {len(config)} configuration arguments')
        config += ['0', 'nullptr'][len(config) - 2:]
        open_at = source.index('(', close)
        if source[close + 3:open_at].strip():
            raise ValueError(f'{path}: unexpected text between >>> and the argument list of {kernel}')
        end = _matching_paren(source, open_at)
        args = source[open_at + 1:end]
        semicolon = end + 1
        while source[semicolon].isspace():
            semicolon += 1
        if source[semicolon] != ';':
            raise ValueError(f'{path}: launch of {kernel} is not a statement')
        original = source[m.start():semicolon + 1]
        flat_args = ' '.join(args.split())
        replacement = (f'::ds_emu::launch(reinterpret_cast<const void *>(&{kernel}), "{kernel}", dim3({config[0]}), dim3({config[1]}), '
                       f'(size_t)({config[2]}), {config[3]}, [&]() {{ {kernel}({flat_args}); }});')
        out.append(source[at:m.start()])
        out.append(replacement + '\n' * original.count('\n'))
        at = semicolon + 1
        launches += 1
    out.append(source[at:])
    text = ''.join(out)
    text, n_dynamic = DYNAMIC_SMEM.subn(lambda d: f'unsigned char *{d.group(1)} = ::ds_emu::dynamic_smem();', text)
    if '<<<' in text or re.search(r'extern\s+__shared__', text):
        raise ValueError(f'{path}: CUDA syntax left after the translation')
    return f'#line 1 "{path}"\n' + text, launches, n_dynamic


def main(argv):
    src, dst = argv[1], argv[2]
    with open(src) as f:
        text, launches, n_dynamic = translate(f.read(), src)
    with open(dst, 'w') as f:
        f.write(text)
    print(f'{src}: {launches} launches, {n_dynamic} dynamic shared-memory declarations')


if __name__ == '__main__':
    main(sys.argv)

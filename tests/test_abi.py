"""CPU: the C-ABI library builds, loads and exports every symbol include/yolact_b200.h
declares (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'yolact_b200.h')).read()
    return sorted(set(re.findall(r'YB_API[^;(]*?\b(yb_\w+)\s*\(', text)))


def test_header_declares_expected_surface():
    syms = declared_symbols()
    for must in ('yb_detect', 'yb_detect_host', 'yb_hard_nms', 'yb_mask_assemble', 'yb_net_create', 'yb_net_forward',
                 'yb_net_detect_host', 'yb_last_error'):
        assert must in syms


def test_library_builds_and_exports_every_symbol():
    from yolact_minimal_b200 import build, _lib
    path = build.build()
    assert os.path.exists(path)
    raw = ctypes.CDLL(path)
    missing = [s for s in declared_symbols() if not hasattr(raw, s)]
    assert not missing, f'symbols declared in the header but not exported: {missing}'
    assert set(_lib.PROTOTYPES) == set(declared_symbols())
    L = _lib.lib()
    assert L.yb_version() == 100
    assert L.yb_launch_count() == 0


def test_no_oracle_import_in_product():
    pkg = os.path.join(ROOT, 'yolact_minimal_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, f


def test_header_is_plain_c_and_links_without_torch(tmp_path):
    """The boundary is a C ABI: the header must compile as C99 (no C++, no torch types) and a C program must link against
    the library and call the entry points that need no device."""
    import subprocess
    from yolact_minimal_b200 import build
    lib = build.build()
    src = tmp_path / 'demo.c'
    src.write_text('#include <stdio.h>\n#include "yolact_b200.h"\n'
                   'int main(void) {\n'
                   '  yb_detect_params p = {0.05f, 0.5f, 200, 100, 81, 32, 0, 550.0f, 0};\n'
                   '  printf("%d %zu\\n", yb_version(), yb_detect_workspace_bytes(1, 19248, &p));\n'
                   '  return yb_detect(0, 0, 0, 0, 1, 19248, &p, 0, 0, 0, 0, 0, 0, 0, 0, 0) == YB_OK;   /* NULL pointers must be refused */\n'
                   '}\n')
    exe = tmp_path / 'demo'
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src),
                           '-o', str(exe), lib, '-Wl,-rpath,' + os.path.dirname(lib)])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out
    ver, ws = out.stdout.split()
    assert int(ver) == 100 and int(ws) > 0

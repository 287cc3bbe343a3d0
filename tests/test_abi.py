"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol that
include/gnndelete_b200.h declares (no compute calls without a GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, 'include', 'gnndelete_b200.h')


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(gd_[a-z0-9_]+)\s*\(', text)))


def test_header_declares_the_expected_groups():
    syms = declared_symbols()
    for must in ['gd_csr_from_coo', 'gd_spmm', 'gd_gat_fwd', 'gd_gat_bwd_dst', 'gd_gat_bwd_src', 'gd_rgcn_conv',
                 'gd_gemm_rows', 'gd_gemm_tn_rows', 'gd_edge_loss_fwd', 'gd_khop_masks', 'gd_to_undirected',
                 'gd_adam_step']:
        assert must in syms


def test_library_exports_every_declared_symbol(lib):
    from gnndelete_b200 import _lib
    exported = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    names = set(re.findall(r'\b(gd_[a-z0-9_]+)\b', exported))
    missing = [s for s in declared_symbols() if s not in names]
    assert not missing, f'declared in the header but not exported: {missing}'
    # and the ctypes table covers exactly the header
    assert sorted(_lib.SIGNATURES) == declared_symbols()


def test_library_targets_sm_100a(lib):
    from gnndelete_b200 import _lib
    out = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert 'sm_100a' in out, out


def test_version_and_error_string(lib):
    assert lib.gd_version() >= 100
    assert isinstance(lib.gd_last_error(), bytes)
    assert lib.gd_launch_count() >= 0
    # pure host-side query functions are callable without a device
    assert lib.gd_csr_workspace_bytes(1000, 100) > 0
    assert lib.gd_gemm_tn_workspace_bytes(10_000, 128, 128) >= 128 * 128 * 4
    assert lib.gd_khop_workspace_bytes(1 << 20) >= 3 * (1 << 20) // 8


def test_no_cpu_fallback():
    """Kernels refuse CPU tensors instead of silently computing on the host."""
    import pytest
    import torch
    from gnndelete_b200 import _lib
    with pytest.raises(RuntimeError):
        _lib.ptr(torch.zeros(4))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'gnndelete_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'framework')):
        for f in files:
            if f.endswith('.py'):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f


def _struct_fields(name):
    """Field names and C types of `typedef struct <tag> { ... } <name>;` in the header, in order."""
    text = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    body = re.search(r'typedef\s+struct\s+\w+\s*\{([^{}]*)\}\s*' + name + r'\s*;', text).group(1)
    out = []
    for decl in body.split(';'):
        decl = decl.strip()
        if decl:
            m = re.match(r'(.*?)(\w+)$', decl, flags=re.S)
            out.append((m.group(2), ' '.join(m.group(1).split())))
    return out


@pytest.mark.parametrize('cname,pyname', [('gd_csr_t', 'CsrStruct'), ('gd_spmm_bplan_t', 'BplanStruct')])
def test_ctypes_structs_match_the_header(cname, pyname):
    """The plain-struct arguments are laid out by hand on the Python side: same fields, order and widths."""
    from gnndelete_b200 import _lib
    want = _struct_fields(cname)
    got = getattr(_lib, pyname)._fields_
    assert [n for n, _ in want] == [n for n, _ in got]
    for (name, ctype), (_, pytype) in zip(want, got):
        if '*' in ctype:
            assert pytype is ctypes.c_void_p, name
        elif ctype == 'int64_t':
            assert pytype is ctypes.c_int64, name
        elif ctype == 'int32_t':
            assert pytype is ctypes.c_int32, name
        else:
            raise AssertionError(f'unhandled C type {ctype!r} for {name}')


def test_ctypes_signatures_match_the_prototypes():
    """Every prototype of the header against the hand-written ctypes table: parameter count and, per parameter,
    pointer / 64-bit / 32-bit / float class (a wrong width in a ctypes call corrupts arguments silently)."""
    from gnndelete_b200 import _lib
    text = re.sub(r'/\*.*?\*/', '', open(HEADER).read(), flags=re.S)
    protos = dict(re.findall(r'\b(gd_[a-z0-9_]+)\s*\(([^()]*)\)\s*;', text))
    assert sorted(protos) == sorted(_lib.SIGNATURES)

    def cls(c_decl):
        c_decl = ' '.join(c_decl.split())
        if c_decl in ('void', ''):
            return None
        if '*' in c_decl or c_decl.startswith('gd_stream_t'):
            return ctypes.c_void_p
        for key, t in (('int64_t', ctypes.c_int64), ('int32_t', ctypes.c_int32), ('size_t', ctypes.c_size_t),
                       ('float', ctypes.c_float)):
            if c_decl.startswith(key + ' ') or c_decl == key:
                return t
        raise AssertionError(f'unhandled parameter {c_decl!r}')

    for name, params in protos.items():
        want = [cls(p) for p in params.split(',')]
        want = [w for w in want if w is not None]
        got = list(_lib.SIGNATURES[name][1])
        got = [ctypes.c_void_p if isinstance(g, type) and issubclass(g, ctypes._Pointer) else g for g in got]
        assert got == want, f'{name}: ctypes {got} vs header {want}'


def test_header_is_plain_c_and_links_from_c(tmp_path, lib):
    """The boundary is a C ABI: the header compiles as C99 and as C++, and a C program links against the library and
    calls its host-side entry points (no device needed)."""
    from gnndelete_b200 import _lib
    for std, lang in (('c99', 'c'), ('c++17', 'c++')):
        r = subprocess.run(['gcc' if lang == 'c' else 'g++', f'-std={std}', '-Wall', '-pedantic', '-fsyntax-only', '-x', lang,
                            HEADER], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    src = tmp_path / 'probe.c'
    src.write_text('#include <stdio.h>\n#include "gnndelete_b200.h"\n'
                   'int main(void) {\n'
                   '    printf("%d %zu %zu\\n", gd_version(), gd_csr_workspace_bytes(1000, 100), gd_row_mse_workspace_bytes(10));\n'
                   '    return gd_last_error() == NULL;\n}\n')
    exe = tmp_path / 'probe'
    libdir = os.path.dirname(_lib.LIB_PATH)
    r = subprocess.run(['gcc', '-std=c99', str(src), '-I', os.path.dirname(HEADER), '-L', libdir, '-lgnndelete_b200',
                        f'-Wl,-rpath,{libdir}', '-o', str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    version, csr_ws, mse_ws = r.stdout.split()
    assert int(version) >= 100 and int(csr_ws) > 0 and int(mse_ws) > 0


def test_argument_errors_are_reported_not_crashed(lib):
    """Error behaviour of the boundary: invalid arguments come back as negative status codes with a message in
    gd_last_error() before anything touches the device."""
    vp = ctypes.c_void_p
    one = vp(16)                                   # any non-NULL address: the calls below must fail before dereferencing it
    rc = lib.gd_row_mse_fwd_bwd(None, 64, one, 64, 64, 10, one, one, 1.0, 1.0, 0.5, 0.5, None, 0, one, one, 1 << 20, None)
    assert rc == -1 and b'gd_row_mse_fwd_bwd' in lib.gd_last_error()
    rc = lib.gd_row_mse_fwd_bwd(one, 32, one, 64, 64, 10, one, one, 1.0, 1.0, 0.5, 0.5, None, 0, one, one, 1 << 20, None)
    assert rc == -1 and b'leading dimension' in lib.gd_last_error()
    rc = lib.gd_row_mse_fwd_bwd(one, 64, one, 64, 64, 10, one, one, 1.0, 1.0, 0.5, 0.5, None, 0, one, one, 8, None)
    assert rc == -3 and b'workspace' in lib.gd_last_error()
    rc = lib.gd_pair_decode(None, 64, 64, one, one, 5, None, None, one, None)
    assert rc == -1 and b'gd_pair_decode' in lib.gd_last_error()
    assert lib.gd_pair_decode(None, 64, 64, None, None, 0, None, None, None, None) == 0      # empty input: nothing to do

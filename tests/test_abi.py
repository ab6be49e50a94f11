"""CPU: the C-ABI library builds/loads and exports every symbol include/nbasr.h declares (no compute calls)."""
import ctypes
import os
import re

from nb_asr_b200 import _build, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'nbasr.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(nbasr_[a-z0-9_]+)\s*\(', src)))


def test_header_symbols_are_exported():
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names          # the ctypes table binds exactly the header
    assert lib.nbasr_version() >= 100


def test_struct_layouts_match_c(tmp_path):
    """Compile the header with gcc and compare sizeof/offsetof with the ctypes mirrors."""
    import subprocess
    prog = tmp_path / 'sz.c'
    prog.write_text('''#include <stdio.h>
#include <stddef.h>
#include "nbasr.h"
int main(void) {
  printf("%zu %zu %zu %zu ", sizeof(nbasr_epilogue), sizeof(nbasr_gemm), sizeof(nbasr_wgrad), sizeof(nbasr_gconv));
  printf("%zu %zu %zu %zu %zu\\n", offsetof(nbasr_epilogue, out), offsetof(nbasr_epilogue, mask_rows),
         offsetof(nbasr_gemm, epi), offsetof(nbasr_gconv, epi), offsetof(nbasr_wgrad, dw));
  return 0;
}''')
    exe = tmp_path / 'sz'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(prog), '-o', str(exe)])
    got = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    E, G, W, GC = _lib.Epilogue, _lib.Gemm, _lib.Wgrad, _lib.GConv
    exp = [ctypes.sizeof(E), ctypes.sizeof(G), ctypes.sizeof(W), ctypes.sizeof(GC), E.out.offset, E.mask_rows.offset,
           G.epi.offset, GC.epi.offset, W.dw.offset]
    assert got == exp


def test_library_is_in_tree_and_sm100a():
    assert os.path.dirname(_build.LIB).endswith(os.path.join('nb_asr_b200', 'csrc'))
    assert 'arch=compute_100a,code=sm_100a' in _build.NVCC_FLAGS

"""pascal/nnCuda.pas (the FreePascal unit that replaces the reference's source/nncuda.pas) is generated from
include/tensorium_b200.h by tools/gen_pascal.py.  FreePascal is not installed here, so the unit is checked structurally:
it is up to date, binds every exported symbol exactly once with the header's arity, and carries every public method of
the reference class TNNCuda<T> (source/nncuda.pas:100-158)."""
import importlib.util
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def gen():
    spec = importlib.util.spec_from_file_location("gen_pascal", os.path.join(ROOT, "tools", "gen_pascal.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


# the reference's public operator methods, in declaration order (source/nncuda.pas:100-158)
REFERENCE_METHODS = """deviceCount Create Destroy CompileLog createDeviceBuffer freeDeviceBuffer readBuffer writeBuffer ActivateArray
activateArraySWISH DeriveArray forwardBias backwardBias gemm gemmBatched gemmStridedBatched addvv subvv mulvv fmavv axpy power
scale crossEntropyLogistic fill copy softmaxBatch crossEntropySoftmax forwardMaxPool backwardMaxPool im2col col2im upSample fmavss
meansAndVars means variances normalize meansAndVarsDelta normalizeDelta addDots forwardScale forwardScaleAdd forwardDropout
backwardDropout costL2 clamp inverseSqrt finish compileToCUBIN loadCUBIN compileFile loadCUBinFile""".split()


def test_unit_is_up_to_date_and_binds_every_symbol_once():
    g = gen()
    text = g.generate()   # asserts symbol existence and call arity for every class method
    assert open(os.path.join(ROOT, "pascal", "nnCuda.pas")).read() == text, "run python tools/gen_pascal.py"
    decls = g.parse_header()
    assert len(decls) >= 90
    for ret, name, params in decls:
        hits = re.findall(rf"^(?:function|procedure) {name}\((.*?)\)(?:: \w+)?; cdecl; external TNNLIB;$", text, flags=re.M)
        assert len(hits) == 1, name
        n = len([p for p in hits[0].split(";") if p.strip()])
        assert n == len(params), (name, n, len(params))


def test_every_reference_method_is_declared_and_implemented():
    text = gen().generate()
    iface, impl = text.split("implementation", 1)
    for m in REFERENCE_METHODS:
        assert re.search(rf"\b(?:procedure|function|constructor|destructor)\s+{m}\b", iface), m
        assert re.search(rf"\b(?:procedure|function|constructor|destructor)\s+TNNCuda<T>\.{m}\b", impl), m


def test_header_symbols_are_exported_by_the_library_and_bound_by_python():
    """The same symbol list three ways: the header, the ctypes table of the Python mirror, the Pascal unit."""
    from tensorium_b200 import _lib
    names = {name for _, name, _ in gen().parse_header()}
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)
    for _, name, params in gen().parse_header():
        assert len(_lib.SIGNATURES[name][1]) == len(params), name

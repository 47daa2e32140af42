"""Every switch mpros_ctx_set_option accepts (csrc/ctx.cu: apply_option) is documented where a caller looks for it: the
header comment of mpros_ctx_set_option and the Python handle's docstring; every MPROS_* environment variable the library reads
maps onto one of them. CPU only (text of the sources)."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "hybrid-astar-planner_b200")


def read(*parts):
    return open(os.path.join(*parts)).read()


def test_every_option_is_documented():
    ctx = read(PKG, "csrc", "ctx.cu")
    body = ctx[ctx.index("static bool apply_option"):ctx.index("static void free_plan_scratch")]
    keys = re.findall(r'k == "([a-z0-9_]+)"', body)
    assert len(keys) >= 8 and len(set(keys)) == len(keys)
    hdr = read(ROOT, "include", "mpros_gpu.h")
    doc = hdr[hdr.index("Developer / tuning switches"):hdr.index("MPROS_API int mpros_ctx_set_option")]
    py = read(PKG, "mpros_b200.py")
    pydoc = py[py.index("def set_option"):py.index("def release_workspaces")]
    for k in keys:
        assert '"%s"' % k in doc, "include/mpros_gpu.h does not document option %s" % k
        assert '"%s"' % k in pydoc, "mpros_b200.Context.set_option does not list option %s" % k


def test_environment_is_read_once_and_maps_to_options():
    ctx = read(PKG, "csrc", "ctx.cu")
    env = re.findall(r'apply_option\(c->opt, "([a-z0-9_]+)", std::getenv\("(MPROS_[A-Z0-9_]+)"\)\)', ctx)
    assert len(env) >= 8
    for key, var in env:
        assert var == "MPROS_" + key.upper(), (key, var)
    # nowhere else in the library
    for f in os.listdir(os.path.join(PKG, "csrc")):
        if f == "ctx.cu":
            continue
        assert "getenv" not in read(PKG, "csrc", f), "%s reads the environment (options are read once, in mpros_ctx_create)" % f

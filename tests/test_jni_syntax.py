"""The JVM binding cannot be compiled here (no JDK / scalac).  What CAN be checked is checked: the JNI shim type-checks
against a minimal stand-in for <jni.h> (signatures from the JNI specification) and against include/snappy_gpu.h; the struct
offsets the Scala serializer hard-codes equal the ABI's; no Get*Critical section exists in the shim."""
import ctypes as C
import os
import re
import subprocess

from snappydata_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SHIM = os.path.join(ROOT, "jvm", "native", "snappy_gpu_jni.c")
SCALA = os.path.join(ROOT, "jvm", "src", "main", "scala", "org", "apache", "spark", "sql", "execution", "columnar", "gpu", "GpuPlanSerializer.scala")


def test_jni_shim_type_checks_against_the_abi_header():
    r = subprocess.run(["gcc", "-fsyntax-only", "-Wall", "-Wextra", "-Werror", "-I" + os.path.join(ROOT, "jvm", "native", "mock"),
                        "-I" + os.path.join(ROOT, "include"), SHIM], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_no_critical_sections_and_natives_match_the_scala_declarations():
    src = open(SHIM).read()
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "GetPrimitiveArrayCritical" not in code and "GetStringCritical" not in code
    natives = set(re.findall(r"JFN\((\w+)\)\(JNIEnv", src))
    scala = open(os.path.join(ROOT, "jvm", "src", "main", "scala", "io", "snappydata", "gpu", "SnappyGpuNative.scala")).read()
    declared = set(re.findall(r"@native def (\w+)\(", scala))
    assert natives == declared, (natives ^ declared)


def test_struct_layout_constants_of_the_scala_serializer():
    want = {}
    for st in (capi.sd_column, capi.sd_expr, capi.sd_agg, capi.sd_plan_desc, capi.sd_literal, capi.sd_batch):
        want[st.__name__] = (C.sizeof(st), {n: getattr(st, n).offset for n, _ in st._fields_})
    text = open(os.path.join(ROOT, "jvm", "abi_offsets.txt")).read()
    for name, (size, offs) in want.items():
        line = [l for l in text.splitlines() if l.startswith(name + " ")][0]
        assert f"size {size}:" in line
        for f, o in offs.items():
            assert f"{f}@{o}" in line.split(), (name, f, o)
    scala = open(SCALA).read()
    for const, st in (("SIZEOF_COLUMN", capi.sd_column), ("SIZEOF_EXPR", capi.sd_expr), ("SIZEOF_AGG", capi.sd_agg),
                      ("SIZEOF_DESC", capi.sd_plan_desc), ("SIZEOF_LITERAL", capi.sd_literal)):
        assert re.search(rf"{const} = {C.sizeof(st)}\b", scala), const
    assert re.search(rf"VERSION = {capi.SD_ABI_VERSION}\b", scala)
    # the offsets `write` / `writeLiterals` poke: sd_plan_desc fields and sd_literal fields
    d = want["sd_plan_desc"][1]
    for field, off in (("ncols", 4), ("cols", 8), ("nexprs", 16), ("exprs", 24), ("filter", 32), ("nkeys", 36), ("keys", 40), ("naggs", 48),
                       ("aggs", 56), ("nproj", 64), ("proj", 72), ("nliterals", 80), ("literal_types", 88), ("flags", 96)):
        assert d[field] == off
    lit = want["sd_literal"][1]
    assert (lit["is_null"], lit["i"], lit["d"], lit["s"], lit["slen"]) == (4, 8, 16, 24, 32)


def test_every_scala_native_has_a_jni_function_and_vice_versa():
    """SnappyGpuNative.scala's @native declarations and snappy_gpu_jni.c's JFN(...) entry points name the same set (a missing
    one is an UnsatisfiedLinkError at the first call on a box that does have a JVM)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    scala = open(os.path.join(root, "jvm/src/main/scala/io/snappydata/gpu/SnappyGpuNative.scala")).read()
    c = open(os.path.join(root, "jvm/native/snappy_gpu_jni.c")).read()
    declared = set(re.findall(r"@native\s+def\s+(\w+)", scala))
    defined = set(re.findall(r"JNICALL\s+JFN\((\w+)\)", c))
    assert declared == defined, (sorted(declared - defined), sorted(defined - declared))
    assert {"storeCreate", "storePutBatch", "planScanStore", "commCreate", "planExchange", "hostAlloc"} <= declared

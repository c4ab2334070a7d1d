"""The JNI shim (java/) cannot be compiled into a JVM here (no JDK in the image), but it can be held to
account on the CPU box: every `native` method java/embedding/DgeNative.java declares must have a
Java_embedding_DgeNative_<name> body in java/jni/dge_jni.c with the right number of parameters, every such body must be
declared, and the C file must pass gcc's syntax and type check against include/dge.h (java/jni/include/jni.h is a
declaration subset for exactly this purpose).  This is the binding a maintainer adds under LayeredGraph.java:157-252
and DeepWalk.java:73-82."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JAVA = os.path.join(ROOT, "java", "embedding", "DgeNative.java")
GLUE = os.path.join(ROOT, "java", "jni", "dge_jni.c")


def _java_natives():
    src = re.sub(r"/\*.*?\*/", "", open(JAVA).read(), flags=re.S)
    out = {}
    for m in re.finditer(r"public\s+static\s+native\s+[\w\[\]]+\s+(\w+)\s*\(([^)]*)\)\s*;", src):
        args = [a for a in m.group(2).split(",") if a.strip()]
        out[m.group(1)] = len(args)
    return out


def _glue_bodies():
    src = re.sub(r"/\*.*?\*/", "", open(GLUE).read(), flags=re.S)
    out = {}
    for m in re.finditer(r"NAT\(\s*\w+\s*,\s*(\w+)\s*\)\s*\(([^)]*)\)\s*\{", src):
        args = [a for a in m.group(2).split(",") if a.strip()]
        out[m.group(1)] = len(args) - 2            # JNIEnv *, jclass
    return out


def test_every_native_method_has_a_body_and_vice_versa():
    java, glue = _java_natives(), _glue_bodies()
    assert len(java) >= 30
    assert sorted(java) == sorted(glue), (sorted(set(java) - set(glue)), sorted(set(glue) - set(java)))
    for name, n in java.items():
        assert glue[name] == n, (name, n, glue[name])


def test_glue_passes_the_c_compiler():
    r = subprocess.run(["gcc", "-fsyntax-only", "-std=c11", "-Wall", "-Wextra", "-Werror",
                        "-I" + os.path.join(ROOT, "java", "jni", "include"), "-I" + os.path.join(ROOT, "include"), GLUE],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_the_java_example_only_calls_declared_natives():
    java = _java_natives()
    for f in os.listdir(os.path.dirname(JAVA)):
        if f.endswith(".java") and f != "DgeNative.java":
            for name in re.findall(r"DgeNative\.(\w+)\s*\(", open(os.path.join(os.path.dirname(JAVA), f)).read()):
                assert name in java, (f, name)

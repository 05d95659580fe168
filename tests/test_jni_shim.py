"""The JNI boundary as files: jni/mmplace_jni.c must compile against include/mmplace.h (the JDK's jni.h when
JAVA_HOME is set, the compile-check stub jni/stub/jni.h otherwise), reference EVERY entry point the header declares, and
have one JNI function per native the Java class declares (java/.../gpu/MmPlace.java)."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shim_compiles_and_binds_every_entry_point():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "jni"), "-s"])
    hdr = open(os.path.join(ROOT, "include", "mmplace.h")).read()
    declared = set(re.findall(r"\b(mmp_[a-z0-9_]+)\s*\(", hdr))
    obj = os.path.join(ROOT, "jni", "_build", "mmplace_jni.o")
    if os.path.exists(obj):
        und = subprocess.check_output(["nm", "-u", obj], text=True)
    else:  # JAVA_HOME build: the shared object
        und = subprocess.check_output(["nm", "-D", "-u", os.path.join(ROOT, "jni", "libmmplace_jni.so")], text=True)
    used = set(re.findall(r"\b(mmp_[a-z0-9_]+)\b", und))
    assert declared <= used, sorted(declared - used)


def test_every_java_native_has_its_jni_function():
    java = open(os.path.join(ROOT, "java", "com", "ibm", "watson", "modelmesh", "gpu", "MmPlace.java")).read()
    natives = set(re.findall(r"static native [\w\[\]\.]+ (\w+)\(", java))
    shim = open(os.path.join(ROOT, "jni", "mmplace_jni.c")).read()
    funcs = set(re.findall(r"\bFN\((\w+)\)", shim)) - {"name"}
    assert natives == funcs, (sorted(natives - funcs), sorted(funcs - natives))
    assert len(natives) >= 49

"""The Java side of the boundary cannot be compiled here (no JDK), so it is checked as far as a C compiler and a parser go:
the JNI glue compiles against the product header with a stub <jni.h> (signature drift between include/rapid_b200.h and
java/jni/rapid_jni.c fails the build), every `native` of com.vrg.rapid.gpu.Native has exactly one JNI export with the same
number of arguments, and every class INTEGRATION.md names exists with the reference interface it claims to implement."""
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
JNI_C = os.path.join(ROOT, "java", "jni", "rapid_jni.c")
NATIVE = os.path.join(ROOT, "java", "com", "vrg", "rapid", "gpu", "Native.java")
PKG = os.path.join(ROOT, "java", "com", "vrg", "rapid")


def _split_args(s):
    s = s.strip()
    if not s:
        return []
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "<([":
            depth += 1
        elif ch in ">)]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def _java_natives():
    src = re.sub(r"/\*.*?\*/", "", open(NATIVE).read(), flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    found = {}
    for m in re.finditer(r"public\s+static\s+native\s+([\w\[\]]+)\s+(\w+)\s*\((.*?)\)\s*;", src, flags=re.S):
        found[m.group(2)] = (m.group(1), _split_args(m.group(3)))
    return found


def _jni_exports():
    src = re.sub(r"/\*.*?\*/", "", open(JNI_C).read(), flags=re.S)
    found = {}
    for m in re.finditer(r"JNIEXPORT\s+(\w+)\s+JNICALL\s+Java_com_vrg_rapid_gpu_Native_(\w+)\s*\((.*?)\)\s*\{", src, flags=re.S):
        assert m.group(2) not in found, "duplicate export " + m.group(2)
        found[m.group(2)] = (m.group(1), _split_args(m.group(3)))
    return found


JAVA_TO_JNI = {"int": "jint", "long": "jlong", "byte": "jbyte", "boolean": "jboolean", "void": "void", "String": "jstring",
               "int[]": "jintArray", "long[]": "jlongArray", "byte[]": "jbyteArray", "ByteBuffer": "jobject"}


def test_jni_glue_compiles_against_the_product_header():
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    cmd = [gcc, "-std=c11", "-fsyntax-only", "-Wall", "-Wextra", "-Wno-unused-parameter", "-Werror",
           "-I" + os.path.join(ROOT, "tests", "jni_stub"), "-I" + os.path.join(ROOT, "include"), JNI_C]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]


def test_every_native_has_one_export_with_matching_signature():
    nat, exp = _java_natives(), _jni_exports()
    assert len(nat) >= 50
    assert set(nat) == set(exp), "natives without export: %s; exports without native: %s" % (
        sorted(set(nat) - set(exp)), sorted(set(exp) - set(nat)))
    for name, (jret, jargs) in nat.items():
        cret, cargs = exp[name]
        assert JAVA_TO_JNI[jret] == cret, (name, jret, cret)
        assert len(cargs) == len(jargs) + 2, (name, jargs, cargs)          # + JNIEnv*, jclass
        assert cargs[0].startswith("JNIEnv*") and cargs[1].startswith("jclass"), (name, cargs[:2])
        for ja, ca in zip(jargs, cargs[2:]):
            jt = ja.replace("final ", "").split()[0]
            ct = ca.split()[0]
            assert JAVA_TO_JNI[jt] == ct, "%s: Java %r vs C %r" % (name, ja, ca)


def test_every_c_entry_point_the_glue_calls_is_declared():
    hdr = open(os.path.join(ROOT, "include", "rapid_b200.h")).read()
    declared = set(re.findall(r"\b(rapid_\w+)\s*\(", hdr))
    called = set(re.findall(r"\b(rapid_\w+)\s*\(", open(JNI_C).read()))
    assert called <= declared, sorted(called - declared)


@pytest.mark.parametrize("cls,needs", [
    ("GpuMultiNodeCutDetector", ["aggregateForProposal", "invalidateFailingEdges", "getNumProposals", "clear"]),
    ("GpuMembershipView", ["getRing", "getObserversOf", "getSubjectsOf", "getExpectedObserversOf", "getRingNumbers",
                           "isHostPresent", "getCurrentConfigurationId"]),
    ("GpuFastPaxosTally", ["handleFastRoundProposal"]),
    ("GpuSimMessaging", ["implements IMessagingClient, IMessagingServer", "sendMessageBestEffort", "sendMessage", "shutdown",
                         "setMembershipService", "start"]),
    ("ScenarioFailureDetector", ["implements IEdgeFailureDetectorFactory", "createInstance(final Endpoint subject, final Runnable notifier)",
                                 "addFailedNodes"]),
])
def test_seam_classes_exist_with_the_reference_method_names(cls, needs):
    p = os.path.join(PKG, cls + ".java")
    assert os.path.exists(p), p
    src = open(p).read()
    assert "package com.vrg.rapid;" in src
    for n in needs:
        assert n in src, "%s lacks %r" % (cls, n)
    # every Native.* call names a declared native
    nat = _java_natives()
    for m in re.finditer(r"Native\.(\w+)\s*\(", src):
        assert m.group(1) in nat, "%s calls undeclared Native.%s" % (cls, m.group(1))


def test_integration_md_names_only_existing_java_classes():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    have = {f[:-5] for f in os.listdir(PKG) if f.endswith(".java")} | {"Native"}
    for name in set(re.findall(r"\b(Gpu[A-Z]\w+|ScenarioFailureDetector)\b", text)):
        assert name in have, "INTEGRATION.md names %s, which is not under java/" % name

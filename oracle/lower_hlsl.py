#!/usr/bin/env python
"""Lower the DECLARATION syntax of an HLSL compute shader to C++ (stdout).  TEST INFRASTRUCTURE ONLY.

Used by oracle/Makefile to compile the reference's Particles/nBodyGravityCS.hlsl, where it lies
under /root/reference, into oracle/_ref/ with g++ and oracle/hlsl_shim.hpp.  Only declarations are
rewritten; statements -- every arithmetic line of bodyBodyInteraction and CSMain -- pass through
byte for byte (the script asserts that).  Rules:

  inout T name                      ->  T &name
  cbuffer X : register(bN) { ... };  ->  the members as globals
  : register(xN)                    ->  (dropped)
  [numthreads(...)]                 ->  (dropped)
  uint3 DTid : SV_DispatchThreadID  ->  uint3 DTid
"""
import re
import sys


def lower(src: str) -> str:
    out = src
    out = re.sub(r"\binout\s+(\w+)\s+(\w+)", r"\1 &\2", out)
    out = re.sub(r"cbuffer\s+\w+\s*:\s*register\s*\(\s*\w+\s*\)\s*\{(.*?)\}\s*;", r"\1", out, flags=re.S)
    out = re.sub(r"\s*:\s*register\s*\(\s*\w+\s*\)", "", out)
    out = re.sub(r"^\s*\[numthreads\s*\([^)]*\)\]\s*$", "", out, flags=re.M)
    out = re.sub(r"\s*:\s*SV_\w+", "", out)
    return out


def statements(text: str):
    """every assignment statement (all the arithmetic of the shader), whitespace-normalised"""
    return [" ".join(l.split()) for l in text.splitlines() if "=" in l and l.strip().endswith(";")]


if __name__ == "__main__":
    source = open(sys.argv[1], encoding="utf-8", errors="replace").read()
    lowered = lower(source)
    a, b = statements(source), statements(lowered)
    assert a == b, "lowering touched a statement"
    sys.stdout.write(lowered)

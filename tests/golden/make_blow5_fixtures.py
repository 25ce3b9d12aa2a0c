#!/usr/bin/env python3
"""Small BLOW5 fixtures for the device-side record / signal decoders, written by the reference's own
slow5lib (runs only where /root/reference exists): the first 8 records of tests/golden/ecoli/reads.blow5 re-encoded as
  ecoli8_zlib_svbzd.blow5   record compression zlib,  signal compression svb-zd   (slow5tools' default)
  ecoli8_none_svbzd.blow5   record compression none,  signal compression svb-zd
  ecoli8_zlib_exzd.blow5    record compression zlib,  signal compression ex-zd   (slow5lib >= 1.2)
  ecoli8_none_none.blow5    no compression at all
The signals they must decode to are those of the zlib-only original (tests/blow5.py reads that with Python's zlib)."""
import os, shutil, struct, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference"
sys.path.insert(0, os.path.dirname(HERE))
import blow5

src_file = os.path.join(HERE, "ecoli", "reads.blow5")
f = blow5.Blow5(src_file)
d = f.data
hs = struct.unpack("<I", d[64:68])[0]
end = f.records[7][0] + f.records[7][1]
tmp = tempfile.mkdtemp(prefix="slow5build")
small = os.path.join(tmp, "first8.blow5")
open(small, "wb").write(d[:end] + b"5WOLB")
lib = os.path.join(tmp, "slow5lib")
shutil.copytree(os.path.join(REF, "slow5lib"), lib)
subprocess.check_call(["make", "-s", "-C", lib, "lib/libslow5.a"], stdout=subprocess.DEVNULL)
exe = os.path.join(tmp, "recompress")
subprocess.check_call(["gcc", "-O2", "-I", os.path.join(lib, "include"), os.path.join(HERE, "blow5_recompress.c"),
                       os.path.join(lib, "lib", "libslow5.a"), "-lz", "-lm", "-lpthread", "-o", exe])
for rec, sig in (("zlib", "svb-zd"), ("none", "svb-zd"), ("zlib", "ex-zd")):
    out = os.path.join(HERE, "ecoli", "ecoli8_%s_%s.blow5" % (rec, sig.replace("-", "")))
    subprocess.check_call([exe, small, out, rec, sig])
    print(out, os.path.getsize(out))
shutil.rmtree(tmp)

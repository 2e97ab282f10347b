"""TEST INFRASTRUCTURE: writes <dir>/kjv.txt from the oracle's decode of <dir>/kjv.txt.huff.

The reference ships kjv.txt.huff without kjv.txt and its loadTextFile
(framework/huffdata.c:152-165) crashes on the missing file, so the unmodified reference harness
needs a stand-in to compare against.  The decode is the oracle's restatement of simpleDecode;
the result is only written when its SHA-256 is the digest of the reference's own output
pinned in SURVEY 8(c)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, ROOT)
import oracle_lib as O  # noqa: E402

d = sys.argv[1]
src = os.path.join(d, "kjv.txt.huff")
dst = os.path.join(d, "kjv.txt")
if os.path.exists(dst):
    sys.exit(0)
st = O.load_huff(src)
out = O.simple_decode(st)
assert O.sha256(out) == "e4e21579f6360b35e66dc97b67cd732a3f759623e41e4e077bec039eeb79fd0a", "kjv digest"
out.tofile(dst)
print("wrote", dst, out.size, "bytes")

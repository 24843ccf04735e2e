"""Writes tests/golden/poseidon_constants.sha256 from the REFERENCE's constants (src/parameters.rs:20-146). Run in the
build container, where /root/reference exists; the digest pins the regenerated parameters on the GPU box, where it does
not."""
import hashlib
import os
import re

src = open("/root/reference/src/parameters.rs").read()
blk = src[src.index("pub static ref FR"):src.index('"rate" => 2')]
nums = re.findall(r'"(\d{20,})"', blk)
assert len(nums) == 39 * 3 + 9
digest = hashlib.sha256("\n".join(nums).encode()).hexdigest()
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "poseidon_constants.sha256")
open(out, "w").write(digest + "  src/parameters.rs:20-146 (117 round constants + 9 MDS entries, decimal, newline-joined)\n")
print(digest)

"""profiles/r2_sass_decim1.txt: instruction mix + excerpts of K1's SASS (cuobjdump -sass of the built object)."""
import collections, re, subprocess, sys
obj = sys.argv[1] if len(sys.argv) > 1 else "habdec_b200/csrc/build/decim1.o"
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
blocks = re.split(r"\n\s*Function : ", out)
L = ["SASS evidence for the stage-1 decimator K1 (habdec_b200/csrc/decim1.cu), sm_100a: `cuobjdump -sass %s`" % obj,
     "UBLKCP = 1-D TMA bulk copy (cp.async.bulk; `desc[..]` = with the L2 evict-first cache hint), SYNCS = mbarrier arrive / expect_tx / try_wait,",
     "FFMA2 / FADD2 / FMUL2 = packed two-lane FP32 (complex sample x real tap = ONE FFMA2).  No HMMA / UTCMMA: tensor cores are not used, by design.", ""]
for pat, title in ((r"decim1_kernelILi64ELi348ELb0E", "decim1_kernel<64,348,false>: plain K1 (BASELINE configs[3])"),
                   (r"decim1_kernelILi64ELi348ELb1E", "decim1_kernel<64,348,true>: K1 with the fused NCO (BASELINE configs[4])")):
    for b in blocks[1:]:
        name = b.split("\n", 1)[0]
        if not re.search(pat, name):
            continue
        ops = collections.Counter()
        for l in b.split("\n"):
            m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", l)
            if m:
                ops[m.group(1)] += 1
        L.append("== " + title)
        L.append("   " + name)
        L.append("   static instruction mix: " + ", ".join("%s %d" % kv for kv in ops.most_common(24)))
        L.append("   " + "  ".join("%s=%d" % (k, sum(v for o, v in ops.items() if o.startswith(k))) for k in
                                   ("UBLKCP", "SYNCS", "FFMA2", "FADD2", "FMUL2", "LDS", "STS", "LDG", "STG", "DFMA", "HMMA", "UTCMMA")))
        L.append("   excerpts:")
        for key, cap in (("UBLKCP", 6), ("SYNCS", 6), ("FFMA2", 5), ("FMUL2", 3)):
            n = 0
            for l in b.split("\n"):
                if key in l and n < cap:
                    L.append("     " + l.strip()[:150]); n += 1
        L.append("")
print("\n".join(L))

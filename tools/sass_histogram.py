"""Per-kernel SASS opcode histograms of libg2048.so (cuobjdump -sass; no GPU needed) -> profiles/r02_sass_opcodes.txt.
What the listing is evidence for: sm_100a-only cubins; the bulk-copy engine (UBLKCP = cp.async.bulk) in the streaming
kernels; the Threefry instruction mix (SHF.L.W / LOP3 on the ALU pipe, IMAD.IADD / IADD3 adds) and the shared-memory
table loads (LDS.U16 / LDS.U8) in the play kernels; no tensor-core or tensor-memory instructions -- the path has no
contraction (HMMA / UTCMMA / LDTM would be out of place)."""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "2048-ppo-agent_b200" / "libg2048.so"
KEEP = ("play3_kernel", "play2_kernel", "play_record_compact", "policy_step_obs", "policy_step_kernel", "rollout_steps_kernel",
        "expand_obs_tma_kernel", "pack_samples", "gae_scan_kernel", "gae_scan_fix", "gae_flat4_kernel", "gae_flat3_kernel",
        "gae_time_major_kernel", "gae_time_major_ring_kernel", "gather_samples_tile_kernel", "compact_records_kernel",
        "scan_chunk", "normalize_kernel", "embed_boards_kernel", "embed_grad_partial", "chain_kernel", "replay_envs")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    kernels, name = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            kernels[name] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            kernels[name][m.group(1)] += 1
    out = [f"# SASS opcode histograms of {LIB.name} (tools/sass_histogram.py); cubin architectures: {', '.join(archs)}",
           f"# kernels in the library: {len(kernels)}", ""]
    everything = collections.Counter()
    for c in kernels.values():
        everything.update(c)
    special = {k: v for k, v in everything.items() if re.match(r"(UBLKCP|UTMA|UTC|LDTM|STTM|HMMA|HGMMA|SYNCS|ATOM|RED|BAR|LDS|STS|LDG|STG|SHF|LOP3|IMAD|IADD3|PRMT|SHFL)", k)}
    out.append("## whole library, selected opcode families")
    for k, v in sorted(special.items(), key=lambda kv: -kv[1]):
        out.append(f"{v:8d}  {k}")
    out.append("")
    out.append("tensor-core / tensor-memory opcodes (HMMA, HGMMA, UTC*MMA, LDTM, STTM): "
               + str(sum(v for k, v in everything.items() if re.match(r"(HMMA|HGMMA|UTC\w*MMA|LDTM|STTM)", k))) + " (none: the path has no contraction)")
    out.append("")
    for name, c in kernels.items():
        if not any(k in name for k in KEEP):
            continue
        total = sum(c.values())
        out.append(f"## {name[:150]}")
        out.append(f"   {total} instructions; " + ", ".join(f"{k} {v}" for k, v in c.most_common(14)))
        out.append("")
    path = ROOT / "profiles" / "r02_sass_opcodes.txt"
    path.write_text("\n".join(out) + "\n")
    print("wrote", path, len(kernels), "kernels")


if __name__ == "__main__":
    main()

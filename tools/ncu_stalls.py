"""Per-instruction view of an `ncu --set full --import-source on` capture of a tile kernel: which warp
(producer, MMA issuer, epilogue) spends its samples where.  Usage:
  python tools/ncu_stalls.py gpurun_out/prof.ncu-rep > profiles/rNN_name_stalls.txt"""
import csv
import io
import subprocess
import sys


def main(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    print(rows[0][1][:140] if rows and len(rows[0]) > 1 else "?")
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    S = lambda r: int(r[ix["# Samples"]])
    E = lambda r: int(r[ix["Instructions Executed"]])
    tot = sum(S(r) for r in data)
    print(f"total warp samples {tot}, SASS instructions {len(data)}, warp-level instructions executed {sum(E(r) for r in data):,}")

    def top3(r):
        s = {h: int(r[ix[h]]) for h in stalls if int(r[ix[h]]) > 0}
        return dict(sorted(s.items(), key=lambda kv: -kv[1])[:3])

    print("\n== instructions with the most samples")
    for r in sorted(data, key=lambda r: -S(r))[:16]:
        print(f"{S(r):9d} {100 * S(r) / tot:5.1f}%  executed {E(r):14,d}  threads {r[ix['Avg. Threads Executed']]:>4s}  "
              f"{r[ix['Source']].strip()[:58]:58s} {top3(r)}")
    # the per-stage loops: instructions executed once per pipeline stage
    mma = [r for r in data if "UTCOMMA" in r[ix["Source"]] or "UTCIMMA" in r[ix["Source"]]]
    if mma:
        n_mma = max(E(r) for r in mma)
        for name, count in (("MMA issuer (instructions executed as often as one of the 4 MMAs of a stage)", n_mma),):
            sel = [r for r in data if E(r) == count]
            st = sum(S(r) for r in sel)
            print(f"\n== {name}: {len(sel)} SASS instructions per stage, {st} samples")
            for r in sorted(sel, key=lambda r: -S(r))[:10]:
                print(f"   {S(r):7d} {100 * S(r) / max(st, 1):5.1f}%  {r[ix['Source']].strip()[:90]}")
        tma = [r for r in data if "UTMALDG" in r[ix["Source"]]]
        if tma:
            n_tma = max(E(r) for r in tma)
            sel = [r for r in data if E(r) == n_tma]
            st = sum(S(r) for r in sel)
            print(f"\n== TMA producer (both CTAs of a pair): {len(sel)} SASS instructions per stage, {st} samples")
            for r in sorted(sel, key=lambda r: -S(r))[:8]:
                print(f"   {S(r):7d} {100 * S(r) / max(st, 1):5.1f}%  {r[ix['Source']].strip()[:90]}")
    print("\n== tensor / TMA / tensor-memory instructions")
    for r in data:
        src = r[ix["Source"]]
        if any(k in src for k in ("UTCOMMA", "UTCIMMA", "UTCBAR", "UTMALDG", "LDTM")):
            print(f"   {S(r):7d}  executed {E(r):14,d}  {src.strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1])

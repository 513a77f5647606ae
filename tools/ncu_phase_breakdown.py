"""Where a kernel's instructions and stall samples go along its (straight-line) program, phase by phase (read without a GPU).
usage: python tools/ncu_phase_breakdown.py <report.ncu-rep> [annotated_sass.txt]
Walks the SASS in address order (the compile-time programs are straight-line code, so address order is program order),
labels every instruction with the program phase of the source function ncu maps it to (forward transform, key product,
inverse transform, plain terms, rotation sum, ...; leaf helpers such as shoup_mul inherit the phase they are inlined into)
and prints executed warp instructions and stall samples per segment and per phase.  Needs -lineinfo and --import-source on."""
import csv, subprocess, re, sys, json, os
from collections import defaultdict, OrderedDict
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from ncu_source_breakdown import function_ranges, load_report_sources
rep = sys.argv[1]
load_report_sources(rep)
raw = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows = list(csv.reader(raw.splitlines()))
addr = {}
fpath=None; hdr=None
for r in rows:
    if not r: continue
    if r[0]=="File Path": fpath=r[1]; hdr=None; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if hdr is None: continue
    # columns: Line No, Source, Address, Source(sass), ...
    if r[0]!="":
        try: line=int(r[0])
        except: pass
        continue
    a=r[2]
    if not a.startswith("0x"): continue
    d=dict(zip(hdr[4:], r[4:]))
    addr[int(a,16)] = dict(file=fpath, line=line, sass=r[3].strip(), inst=float(d.get("Instructions Executed") or 0), samp=float(d.get("# Samples") or 0),
        stalls={k[6:]:float(v or 0) for k,v in d.items() if k.startswith("stall_") and "Not Issued" not in k})
ranges={}
def fn(f,line):
    if f not in ranges: ranges[f]=function_ranges(f)
    name="?"
    for first,n in ranges[f]:
        if first<=line: name=n
        else: break
    return name
PH = {"op_fwd":"fwd","ct_bfly":"fwd","fwd_g1":"fwd","fwd_g2":"fwd","inv_core":"inv","gs_bfly":"inv","inv_g1":"inv","inv_g2":"inv",
      "mac_key":"mack","mac_key_smem":"mack","mac_var":"macv","mac_var_smem":"macv","op_addp":"addp","op_fin":"fin","op_st":"st","op_ld":"ld",
      "op_rot":"rot","rot_biased":"rot","rot_ld_pair":"rot","rot_st_pair":"rot","rot_ld128":"rot","op_norm":"norm","crt2_mod_q_f64":"crt","fold_flags":"flags","reduce_q_centered_f64":"fin",
      "cta_lockstep":"sync","pp_acquire":"sync","pp_release":"sync", "op_stg":"stg"}
cur="prolog"; seq=[]  # list of [phase, inst, samp, stalls, ninstr]
for a in sorted(addr):
    e=addr[a]; f=fn(e["file"],e["line"])
    ph=PH.get(f)
    if ph and ph!="sync": 
        if ph!=cur:
            cur=ph
    key = cur if PH.get(f)!="sync" else "sync"
    if not seq or seq[-1][0]!=key: seq.append([key,0.0,0.0,defaultdict(float),0])
    seq[-1][1]+=e["inst"]; seq[-1][2]+=e["samp"]; seq[-1][4]+=1
    for k,v in e["stalls"].items(): seq[-1][3][k]+=v
ti=sum(s[1] for s in seq); ts=sum(s[2] for s in seq)
# merge tiny segments into neighbours for readability
print(f"total inst {ti:.0f} samples {ts:.0f}")
agg=defaultdict(lambda:[0.0,0.0,defaultdict(float)])
for ph,i,s,st,n in seq:
    if i/ti>0.004 or s/ts>0.004:
        top=", ".join(f"{k} {100*v/max(s,1):.0f}" for k,v in sorted(st.items(), key=lambda kv:-kv[1])[:4])
        print(f"  {ph:8s} sass {n:5d}  inst {100*i/ti:6.2f}%  samples {100*s/ts:6.2f}%   {top}")
    agg[ph][0]+=i; agg[ph][1]+=s
    for k,v in st.items(): agg[ph][2][k]+=v
print("--- aggregated by phase")
for ph,(i,s,st) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    top=", ".join(f"{k} {100*v/max(s,1):.0f}" for k,v in sorted(st.items(), key=lambda kv:-kv[1])[:5])
    print(f"  {ph:8s} inst {100*i/ti:6.2f}%  samples {100*s/ts:6.2f}%   {top}")
if len(sys.argv)>2:
    with open(sys.argv[2],"w") as f:
        cur="prolog"
        for a in sorted(addr):
            e=addr[a]; fnn=fn(e["file"],e["line"])
            f.write(f"{a&0xfffff:06x} {fnn:22s} L{e['line']:<5d} {e['samp']:5.0f} {e['sass']}\n")

import collections, csv, io, os, re, subprocess, sys
CSRC='/root/repo/neuro_genetic_pong_self_play_b200/csrc'
cubin='/tmp/core.cubin'
subprocess.check_call(["nvcc","-gencode","arch=compute_100a,code=sm_100a","-lineinfo","-O3","-std=c++17","-cubin","-o",cubin,os.path.join(CSRC,'ngp_core.cu')],stderr=subprocess.DEVNULL)
sass=subprocess.run(["nvdisasm","-g","-c",cubin],capture_output=True,text=True).stdout.split("\n")
rows=list(csv.reader(open(sys.argv[1])))
hdr,data=rows[1],rows[2:]
iE,iS,iT=hdr.index("Instructions Executed"),hdr.index("# Samples"),hdr.index("Thread Instructions Executed")
starts=[i for i,l in enumerate(sass) if l.startswith(".text.") and 'rollout_kernel' in l]
best=None
for st in starts:
    end=next(i for i in range(st+1,len(sass)) if sass[i].startswith("//---------------------") or i==len(sass)-1)
    cur,seq=("?",0),[]
    for l in sass[st:end]:
        m=re.search(r'//## File "([^"]+)", line (\d+)',l)
        if m: cur=(os.path.basename(m.group(1)),int(m.group(2))); continue
        m=re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);",l)
        if m: seq.append((cur,m.group(2)))
    if len(seq)==len(data): best=seq
assert best
frames=float(sys.argv[2])
W=collections.Counter(); T=collections.Counter(); S=collections.Counter()
for (loc,ins),row in zip(best,data):
    W[loc]+=int(row[iE]); T[loc]+=int(row[iT]); S[loc]+=int(row[iS])
tw=sum(W.values()); tt=sum(T.values()); ts=sum(S.values())
print('warp inst/frame',tw/frames,'avg active',tt/tw)
# wasted warp-instructions relative to perfect 32-lane execution of same thread work
waste={k: W[k]-T[k]/32 for k in W}
cache={}
def text(f,ln):
    path=os.path.join(CSRC,f) if os.path.exists(os.path.join(CSRC,f)) else os.path.join(CSRC,'generated',f)
    if path not in cache: cache[path]=open(path).read().split('\n') if os.path.exists(path) else []
    return cache[path][ln-1].strip()[:80] if 0<ln<=len(cache[path]) else ''
print('total wasted/frame',sum(waste.values())/frames)
byfile=collections.Counter()
for k,v in waste.items(): byfile[k[0]]+=v
print({k: round(v/frames) for k,v in byfile.most_common(8)})
for (f,ln),v in sorted(waste.items(), key=lambda kv:-kv[1])[:45]:
    print(f"{v/frames:7.1f} wasted  {W[(f,ln)]/frames:7.1f} w-inst  act {T[(f,ln)]/max(1,W[(f,ln)]):5.1f}  samp {100*S[(f,ln)]/ts:4.1f}%  {f}:{ln} : {text(f,ln)}")

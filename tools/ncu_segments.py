"""One ncu capture cut at the block barriers: for every barrier-to-barrier stretch of the kernel's SASS (address order)
the warp instructions executed, the stall samples, how often the closing barrier ran and the source lines with the most
samples.  TEST / PROFILING TOOL.

    python tools/ncu_segments.py <report.ncu-rep> <library.so>

Reading guide: a warp that waits at a barrier is sampled at the instruction AFTER it, so barrier waits show up as a short
stretch with many samples and few instructions right behind the barrier; the `bar execs` column identifies the stretch
(e.g. warps x iterations for the closing barrier of the frame loop, warps x general frames for the barriers of the
general step).  SASS order is not execution order: a stretch holds whatever code the compiler placed between two barriers.
"""
import csv, io, re, subprocess, os, sys, tempfile, collections
rep, lib = sys.argv[1], os.path.abspath(sys.argv[2])
lines = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout.splitlines()
kernel = next(csv.reader([lines[0]]))[1]
rows=list(csv.DictReader(io.StringIO("\n".join(lines[1:]))))
base=int(rows[0]['Address'],16)
def norm(name):
    name = re.sub(r"\((?:bool|int|unsigned int|long)\)", "", name).replace(" ", "")
    return name.replace("false", "0").replace("true", "1")
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump","-xelf","all",lib],cwd=tmp,stdout=subprocess.DEVNULL,check=True)
    sass=""
    for f in os.listdir(tmp):
        if f.endswith(".cubin"):
            sass+=subprocess.run(["nvdisasm","-g","-c",os.path.join(tmp,f)],stdout=subprocess.PIPE,stderr=subprocess.DEVNULL,text=True).stdout
funcs,cur={},None
for ln in sass.splitlines():
    m=re.match(r"\s*\.text\.(\S+):",ln) or re.match(r"//-+ \.text\.(\S+) -+",ln)
    if m: cur=funcs.setdefault(m.group(1),[]); continue
    if cur is not None: cur.append(ln)
want=None
for mg in funcs:
    dem=subprocess.run(["c++filt",mg],stdout=subprocess.PIPE,text=True).stdout.strip()
    if norm(dem)==norm(kernel): want=mg
off2line={}; where="?"
for ln in funcs[want]:
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',ln)
    if m:
        where="%s:%s"%(os.path.basename(m.group(1)).replace('b2c_',''),m.group(2))
        continue
    m=re.match(r"\s*/\*([0-9a-f]{4,})\*/",ln)
    if m: off2line[int(m.group(1),16)]=where
tot=sum(int(r['# Samples'] or 0) for r in rows)
seg=[]; cur={'n':0,'inst':0,'lines':collections.Counter(),'start':0,'bar':0,'maxinst':0}
for idx,r in enumerate(rows):
    n=int(r['# Samples'] or 0); ie=int(r['Instructions Executed'] or 0)
    w=off2line.get(int(r['Address'],16)-base,'?')
    isbar='BAR.SYNC' in r['Source'] or 'BAR.RED' in r['Source']
    if isbar:
        cur['bar']=n; cur['barinst']=ie; cur['end']=idx
        seg.append(cur); cur={'n':0,'inst':0,'lines':collections.Counter(),'start':idx+1,'bar':0,'maxinst':0}
        continue
    cur['n']+=n; cur['inst']+=ie; cur['lines'][w]+=n; cur['maxinst']=max(cur['maxinst'],ie)
cur['end']=len(rows); seg.append(cur)
print("total samples",tot,"segments",len(seg))
for s in sorted(seg,key=lambda s:-(s['n']+s['bar']))[:45]:
    ls=", ".join("%s(%d)"%(k,v) for k,v in s['lines'].most_common(4))
    print("%5.2f%% work + %5.2f%% bar | rows %5d-%5d | inst %9d | bar execs %8d | %s"%(100*s['n']/tot,100*s['bar']/tot,s['start'],s['end'],s['inst'],s.get('barinst',0),ls))

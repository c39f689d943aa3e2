"""Prototype: entry-level task DAG for LU factor + fwd + bwd; evaluate round/step counts."""
import ctypes as C, numpy as np, sys
sys.path.insert(0,'.')
from xictools_b200 import engine
from xictools_b200.deck import ring_oscillator

def get(deck):
    L=engine.load_library()
    cd,keep=deck.to_c()
    out=(C.c_longlong*32)()
    assert L.b200_deck_analysis_stats(C.byref(cd),out,32)==0
    n,nnz=out[0],out[1]
    Ap=np.zeros(n+1,np.int32);Ai=np.zeros(nnz,np.int32);rp=np.zeros(n,np.int32);cp=np.zeros(n,np.int32)
    L.b200_deck_analysis_arrays.argtypes=[C.c_void_p]*5
    assert L.b200_deck_analysis_arrays(C.byref(cd),Ap.ctypes.data,Ai.ctypes.data,rp.ctypes.data,cp.ctypes.data)==0
    return n,Ap,Ai,rp,cp

def symbolic(n,Ap,Ai,rp,cp):
    rinv=np.zeros(n,int); rinv[rp]=np.arange(n)
    Lcols=[];Ucols=[]
    for j in range(n):
        cj=cp[j]; mark=set(rinv[Ai[Ap[cj]:Ap[cj+1]]])
        for k in range(j):
            if k in mark:
                mark.update(Lcols[k])
        # need ascending processing: marks added with larger index are handled since k ascending
        # (redo properly)
        mark=set(rinv[Ai[Ap[cj]:Ap[cj+1]]])
        k=0
        import heapq
        h=[x for x in mark if x<j]; heapq.heapify(h); seen=set(h)
        U=[]
        while h:
            k=heapq.heappop(h); U.append(k)
            for i in Lcols[k]:
                if i not in mark:
                    mark.add(i)
                    if i<j and i not in seen: heapq.heappush(h,i); seen.add(i)
        Ucols.append(sorted(U)); Lcols.append(sorted(x for x in mark if x>j))
    return Lcols,Ucols

def build(n,Lcols,Ucols,SPLIT=48,CH=32):
    # tasks: dict id -> (ops list of (a_id,b_id), div_id or None); ids: ('e',i,j) entries, ('y',i) fwd, ('x',i) bwd
    tasks={}
    Lset=[set(c) for c in Lcols]
    for j in range(n):
        ops={}
        for k in Ucols[j]:
            for i in Lcols[k]:
                ops.setdefault(i,[]).append((('e',i,k),('e',k,j)))
        for i in Ucols[j]+[j]+Lcols[j]:
            o=ops.get(i,[])
            div=('e',j,j) if i>j else None
            if o or div: tasks[('e',i,j)]=(o,div)
    rows=[[] for _ in range(n)]
    for k in range(n):
        for i in Lcols[k]: rows[i].append((('e',i,k),('y',k)))
    for i in range(n):
        if rows[i]: tasks[('y',i)]=(rows[i],None)
    urows=[[] for _ in range(n)]
    for j in range(n-1,-1,-1):
        for i in reversed(Ucols[j]): urows[i].append((('e',i,j),('x',j)))
    for i in range(n):
        tasks[('x',i)]=(urows[i],('e',i,i))
    # split long tasks
    out={}
    nscratch=0
    for tid,(o,div) in tasks.items():
        if len(o)>SPLIT:
            parts=[]
            for c in range(0,len(o),CH):
                pid=('p',nscratch); nscratch+=1
                out[pid]=(o[c:c+CH],None); parts.append(pid)
            out[tid]=([(p,('neg1',)) for p in parts],div)
        else: out[tid]=(o,div)
    return out,nscratch

def levels(tasks):
    lev={}
    def dep_level(x):
        if x[0]=='x':   # bwd x depends on y
            pass
        return lev.get(x,None)
    # topological by iterative relaxation (ids have natural order but do generic)
    import functools
    sys.setrecursionlimit(100000)
    def L(t):
        if t in lev: return lev[t]
        if t not in tasks:
            # leaf value (assembled in place) ; but ('x',i) initial acc is y_i, ('y',i) w/o task is b_i
            lev[t]=-1; return -1
        o,div=tasks[t]
        m=-1
        for a,b in o:
            m=max(m,L(a),L(b))
        if div: m=max(m,L(div))
        if t[0]=='x': m=max(m,L(('y',t[1])))   # initial accumulator
        lev[t]=m+1
        return m+1
    for t in tasks: L(t)
    return lev

def pack(tasks,lev,lanes):
    nl=max(lev[t] for t in tasks)+1
    by=[[] for _ in range(nl)]
    for t in tasks: by[lev[t]].append(t)
    steps=0;rounds=0;padded=0;padded_active=0
    per_level=[]
    for l in range(nl):
        v=sorted(by[l],key=lambda t:-len(tasks[t][0]))
        ls=0
        for off in range(0,len(v),lanes):
            grp=v[off:off+lanes]
            ns=len(tasks[grp[0]][0])
            steps+=ns;rounds+=1;padded+=ns*lanes;padded_active+=ns*len(grp);ls+=ns
        per_level.append((len(v),ls))
    return nl,rounds,steps,padded,padded_active,per_level

if __name__=='__main__':
    ns=int(sys.argv[1]) if len(sys.argv)>1 else 501
    n,Ap,Ai,rp,cp=get(ring_oscillator(ns))
    Lc,Uc=symbolic(n,Ap,Ai,rp,cp)
    print('n',n,'nnzL',sum(map(len,Lc)),'nnzU',sum(map(len,Uc)))
    for SPLIT,CH in ((10**9,32),(48,32),(24,16)):
        tasks,nscr=build(n,Lc,Uc,SPLIT,CH)
        lev=levels(tasks)
        nops=sum(len(o) for o,d in tasks.values())
        print('SPLIT',SPLIT,'tasks',len(tasks),'ops',nops,'scratch',nscr)
        for lanes in (32,64,128):
            nl,rounds,steps,padded,pa,pl=pack(tasks,lev,lanes)
            print('  lanes',lanes,'levels',nl,'rounds',rounds,'steps',steps,'padded ops',padded,'active-padded',pa)
        print('  per level (ntasks, steps) @128:',pl)

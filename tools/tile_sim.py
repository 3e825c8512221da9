"""Developer tool (CPU): where does the tile padding of the scale-s 3^3 tile book come from?  Rebuilds the neighbour masks of
the benchmark building in numpy and evaluates tile-cut / ordering strategies (fixed 128-row tiles on several sort keys,
optimal cut points by dynamic programming, a two-level scheme, 64- and 32-row tiles).  Results: profiles/experiments/README.md."""
import sys, numpy as np, time
sys.path.insert(0,'/root/repo')
import bench
s=int(sys.argv[1]) if len(sys.argv)>1 else 1
xyz=bench.building(300000)
a=xyz*(50/2**s); a-=a.min(0)
c=np.unique(a.astype(np.int64),axis=0)
n=len(c); print('sites',n)
key=(c[:,0]<<40)|(c[:,1]<<20)|c[:,2]
order=np.argsort(key); keys=key[order]
masks=np.zeros(n,dtype=np.uint32); cnt=np.zeros(27,dtype=np.int64)
k=0
for dx in (-1,0,1):
  for dy in (-1,0,1):
    for dz in (-1,0,1):
      q=((c[:,0]+dx)<<40)|((c[:,1]+dy)<<20)|(c[:,2]+dz)
      pos=np.searchsorted(keys,q); pos[pos>=n]=n-1
      hit=keys[pos]==q
      masks|=(hit.astype(np.uint32)<<k); cnt[k]=hit.sum(); k+=1
pairs=cnt.sum(); print('pairs',pairs)
# sort key: bits ordered by pair count, rarest most significant
rank=np.argsort(np.argsort(-cnt))   # most common -> rank 0 (LSB)
km=np.zeros(n,dtype=np.uint32)
for k in range(27): km|=(((masks>>k)&1)<<rank[k]).astype(np.uint32)
o=np.argsort(km,kind='stable'); ms=masks[o]; kms=km[o]
pc=np.array([bin(x).count('1') for x in range(1<<9)],dtype=np.int32)
def popc(x): return pc[x&511]+pc[(x>>9)&511]+pc[(x>>18)&511]
def cost_fixed(ms,T=128):
    nt=(len(ms)+T-1)//T; pad=np.zeros(nt*T,dtype=np.uint32); pad[:len(ms)]=ms
    u=np.bitwise_or.reduce(pad.reshape(nt,T),axis=1); return int(popc(u).sum())*T
base=cost_fixed(ms); print('fixed tiles: tile_rows/pairs',base/pairs)
# DP over sorted rows: tile = consecutive rows [i,j), j-i<=128, cost popc(OR)*128
N=len(ms)
t0=time.time()
INF=1<<60
best=np.full(N+1,INF,dtype=np.int64); best[0]=0
# forward DP: from i, extend j up to 128, maintaining OR (vectorised over i in chunks is hard; do python loop with numpy inner)
msl=ms.astype(np.int64)
# restrict candidate cuts: cuts only at positions where mask changes, or full 128
chg=np.flatnonzero(np.r_[True,kms[1:]!=kms[:-1],True])  # group starts + N
is_cut=np.zeros(N+1,bool); is_cut[chg]=True
for i in range(N):
    if best[i]>=INF: continue
    seg=msl[i:i+128]
    u=np.bitwise_or.accumulate(seg)
    c128=(popc(u.astype(np.uint32))*128).astype(np.int64)+best[i]
    js=np.arange(i+1,i+len(seg)+1)
    ok=is_cut[js].copy(); ok[-1]=True   # may always take the full 128
    jj=js[ok]; cc=c128[ok]
    np.minimum.at(best,jj,cc)
print('DP (cuts at group boundaries or 128): ratio',best[N]/pairs,'time',time.time()-t0)
print('distinct masks',len(np.unique(masks)))
um,uc=np.unique(masks,return_counts=True)
print('groups >=128:',(uc>=128).sum(),'rows in them',uc[uc>=128].sum(),'of',n,'; groups<16:',(uc<16).sum(),'rows',uc[uc<16].sum())
def ratio_for(order_key):
    o=np.argsort(order_key,kind='stable'); return cost_fixed(masks[o])/pairs
pcs=popc(masks)
print('popcount-major then key:',ratio_for((pcs.astype(np.uint64)<<32)|km))
# common offsets most significant
rank2=np.argsort(np.argsort(cnt))
km2=np.zeros(n,dtype=np.uint32)
for k in range(27): km2|=(((masks>>k)&1)<<rank2[k]).astype(np.uint32)
print('commonest-MSB:',ratio_for(km2))
print('plain mask:',ratio_for(masks))
# lower bound given group structure: sum over groups ceil(size/128)*128*popc ... (each group alone)
lb=sum(((c+127)//128)*128*int(popc(np.uint32(m))) for m,c in zip(um,uc))
print('every group tiled alone (no mixing) ratio',lb/pairs)
# M=64 tiles
def cost_T(ms,T):
    nt=(len(ms)+T-1)//T; pad=np.zeros(nt*T,dtype=np.uint32); pad[:len(ms)]=ms
    u=np.bitwise_or.reduce(pad.reshape(nt,T),axis=1); return int(popc(u).sum())*T
print('T=64 tiles ratio',cost_T(ms,64)/pairs,' T=32',cost_T(ms,32)/pairs)
# two-level: full tiles of big groups alone, leftovers pooled and sorted by key
full=0; left_masks=[]; left_keys=[]
o=np.argsort(km,kind='stable'); ms=masks[o]; ks=km[o]
b=np.flatnonzero(np.r_[True,ks[1:]!=ks[:-1],True])
for i in range(len(b)-1):
    sz=b[i+1]-b[i]; m=ms[b[i]]; nf=sz//128
    full+=nf*128*int(popc(np.uint32(m)))
    r=sz-nf*128
    if r: left_masks.append(np.full(r,m,dtype=np.uint32)); left_keys.append(np.full(r,ks[b[i]],dtype=np.uint32))
lm=np.concatenate(left_masks); lk=np.concatenate(left_keys)
print('leftover rows',len(lm),'full-tile rows cost ratio',full/pairs)
print('two-level ratio',(full+cost_fixed(lm))/pairs)
# leftover pool sorted by popcount-major
lp=popc(lm)
oo=np.argsort((lp.astype(np.uint64)<<32)|lk,kind='stable')
print('two-level, leftovers by popcount-major:',(full+cost_fixed(lm[oo]))/pairs)
print('leftovers T=64:',(full+cost_T(lm,64))/pairs, 'T=32:',(full+cost_T(lm,32))/pairs)

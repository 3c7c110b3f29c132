import zlib, random
P=0xEDB88320
T0=[]
for i in range(256):
    c=i
    for k in range(8): c=(c>>1)^P if c&1 else c>>1
    T0.append(c)
def Z(c,m):
    for _ in range(m): c=T0[c&0xFF]^(c>>8)
    return c
def multmodp(a,b):
    p=0
    for i in range(31,-1,-1):
        if (a>>i)&1: p^=b
        b=(b>>1)^P if b&1 else b>>1
    return p
def xpow(bits):   # x^bits mod P reflected: x^0 = 1<<31
    r=1<<31; base=1<<30; e=bits
    while e:
        if e&1: r=multmodp(base,r)
        base=multmodp(base,base); e>>=1
    return r
# order check
print("x^(2^32-1)==1:", xpow(2**32-1)==(1<<31))
assert Z(0x12345678,7)==multmodp(xpow(56),0x12345678)
XINV512=xpow((2**32-1)-8*512)
assert multmodp(XINV512,xpow(8*512))==(1<<31)
def tables(m):
    return [[Z(b<<(8*j),m) for b in range(256)] for j in range(4)]
U=tables(512); ZT={m:tables(m) for m in (4,8,16,32,64,128,256)}
def app(T,c): return T[0][c&0xFF]^T[1][(c>>8)&0xFF]^T[2][(c>>16)&0xFF]^T[3][c>>24]
def span_raw_X(rows):   # rows: list of 512-byte rows (bytes). returns X such that raw(rows)=Z_4(X)
    K=len(rows)
    V=[0]*128
    for k,row in enumerate(rows):
        for s in range(128):
            w=int.from_bytes(row[4*s:4*s+4],'little')
            x=V[s]^w
            V[s]=app(U,x) if k<K-1 else x
    lvl=V; m=4
    while len(lvl)>1:
        lvl=[app(ZT[m],lvl[2*i])^lvl[2*i+1] for i in range(len(lvl)//2)]
        m*=2
    return lvl[0]
def raw(data):
    c=0
    for d in data: c=T0[(c^d)&0xFF]^(c>>8)
    return c
rng=random.Random(1)
for trial in range(20):
    n=rng.choice([0,1,5,511,512,513,2000,5000,70000])
    mis=rng.randrange(16)
    data=bytes(rng.getrandbits(8) for _ in range(n))
    nv=n+mis
    NV=(nv+511)//512*512
    virt=bytes(mis)+data+bytes(NV-nv)
    span_rows=rng.choice([1,2,3,7,64])
    acc=0
    nrows=NV//512
    for k0 in range(0,nrows,span_rows):
        rows=[virt[512*r:512*r+512] for r in range(k0,min(nrows,k0+span_rows))]
        X=span_raw_X(rows)
        E=min(nrows,k0+span_rows)*512
        e=nv-E+4+512
        assert e>0
        acc^=multmodp(xpow(8*e),X)
    r=multmodp(XINV512,acc)
    assert r==raw(data),(n,mis,span_rows)
    init=rng.getrandbits(32) if trial%2 else 0
    crc=r^multmodp(xpow(8*n),init^0xFFFFFFFF)^0xFFFFFFFF if n else init
    assert crc==zlib.crc32(data,init),(n,init)
print("crc model ok")

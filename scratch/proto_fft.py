import numpy as np
def factorize(n):
    f=[]; 
    for p in (4,2,3,5):
        while n%p==0 and (p!=4 or n%4==0): f.append(p); n//=p
    p=7
    while n>1:
        while n%p==0: f.append(p); n//=p
        p+=2
    return f
def stockham(x, sign=-1, order=None):
    N=len(x); fac=order or sorted(factorize(N), reverse=True)
    W=np.exp(sign*2j*np.pi*np.arange(N)/N)
    a=x.astype(complex).copy(); Ns=1
    for R in fac:
        b=np.empty_like(a); M=N//R
        for j in range(M):
            k=j%Ns
            for q in range(R):
                acc=0
                for r in range(R):
                    idx=(r*(k*(N//(Ns*R)) + q*(N//R)))%N
                    acc+=a[j+r*M]*W[idx]
                b[(j//Ns)*Ns*R + k + q*Ns]=acc
        a=b; Ns*=R
    return a
for N in (214, 30, 107, 64, 1500//4, 314):
    x=np.random.randn(N)+1j*np.random.randn(N)
    print(N, sorted(factorize(N),reverse=True), np.abs(stockham(x)-np.fft.fft(x)).max(), np.abs(stockham(x,+1)-np.fft.ifft(x)*N).max())
# packed two-real FFT cross-power
Sh,Sw=12,10
a=np.random.rand(Sh,Sw); b=np.random.rand(Sh,Sw)
Z=np.fft.fft2(a+1j*b)
Zm=np.conj(np.roll(np.roll(Z[::-1,::-1],1,0),1,1))   # conj(Z(-k))
A=(Z+Zm)/2; B=(Z-Zm)/(2j)
print(np.abs(A-np.fft.fft2(a)).max(), np.abs(B-np.fft.fft2(b)).max())
P=A*np.conj(B); R=P/np.maximum(np.abs(P),1e-14)
# inverse: column ifft then pack row pairs
Y=np.fft.ifft(R,axis=0)*Sh
cc=np.fft.ifft2(R)
out=np.zeros((Sh,Sw))
for y in range(0,Sh,2):
    V=Y[y]+1j*Y[y+1]
    v=np.fft.ifft(V)*Sw
    out[y]=v.real; out[y+1]=v.imag
print(np.abs(out/(Sh*Sw)-cc.real).max(), np.abs(cc.imag).max())

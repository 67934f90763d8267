import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/tmp/w')
import numpy as np
import fhn_sym as M

class Proto:
    """Structured (compressed-Jacobian) algorithm, single chain, FHN noiseless."""
    def __init__(s, T,S,R,y,dl):
        s.T,s.S,s.R,s.y,s.dl=T,S,R,np.asarray(y),dl
        s.parts=[]
        for init in (R, R//2):
            nfull,nrem=divmod(T-init,R)
            nmid = nfull-1 if nrem==0 else nfull
            fin = R if nrem==0 else nrem
            sizes=[init]+[R]*nmid+[fin]
            s.parts.append(sizes)
    def blocks(s,part):
        o=0
        for b,n in enumerate(s.parts[part]):
            yield b,o,n,(b==0),(b==len(s.parts[part])-1)
            o+=n
    def unpack(s,q):
        u=q[:4]; v0=q[4:6]; v=q[6:].reshape(-1,2)
        z=np.array([np.exp(u[0]),np.exp(u[1]),np.exp(u[2]),u[3]])
        dz=np.array([z[0],z[1],z[2],1.0])  # diag of dz/du
        return u,v0,v,z,dz
    def constr(s,q,xobs,part):
        u,v0,v,z,dz=s.unpack(q); S=s.S; out=[]
        for b,o,n,ini,fin in s.blocks(part):
            x = (v0-np.array([0,z[3]])) if ini else xobs[o-1]
            for k in range(n):
                for t in range(S):
                    x=M.step(*z,*x,*v[(o+k)*S+t],s.dl)[:,0]
                if fin or k<n-1: out.append(x[0]-s.y[o+k,0])
                else: out+= [x[0]-xobs[o+k,0], x[1]-xobs[o+k,1]]
        return np.array(out)
    def linearize(s,q,xobs,part):
        """returns dict of per-block compressed jacobian info + chol etc"""
        u,v0,v,z,dz=s.unpack(q); S=s.S; dl=s.dl
        N=s.T*S
        xs=np.zeros((N,2)); K=np.zeros((N,2,2)); Psib=np.zeros((s.T,2,2)); Q=np.zeros((s.T,2,2)); Zt=np.zeros((s.T,2,4))
        B=[]
        for b,o,n,ini,fin in s.blocks(part):
            x = (v0-np.array([0,z[3]])) if ini else xobs[o-1]
            for k in range(n):
                g0=(o+k)*S
                for t in range(S):
                    xs[g0+t]=x
                    x=M.step(*z,*x,*v[g0+t],dl)[:,0]
                Psi=np.eye(2)
                for t in reversed(range(S)):
                    Bt=M.Fv(*z,*xs[g0+t],*v[g0+t],dl); Gt=M.Fz(*z,*xs[g0+t],*v[g0+t],dl); Ft=M.Fx(*z,*xs[g0+t],*v[g0+t],dl)
                    K[g0+t]=Psi@Bt; Q[o+k]+=K[g0+t]@K[g0+t].T; Zt[o+k]+=Psi@Gt
                    Psi=Psi@Ft
                Psib[o+k]=Psi
            # obs level
            rows=[]  # list of (k, Hrow)
            for k in range(n):
                if fin or k<n-1: rows.append((k,np.array([1.0,0.0])))
                else: rows+= [(k,np.array([1.0,0.0])),(k,np.array([0.0,1.0]))]
            Su=np.zeros((2,4)); P=np.zeros((2,2))
            if ini:
                Su[1,3]=-1.0*dz[3]; P=np.eye(2)
            Sus=[];Ps=[]
            for k in range(n):
                Su=Psib[o+k]@Su+Zt[o+k]*dz[None,:]
                P=Psib[o+k]@P@Psib[o+k].T+Q[o+k]
                Sus.append(Su);Ps.append(P)
            nr=len(rows); A=np.zeros((nr,4)); D=np.zeros((nr,nr))
            for i,(k,h) in enumerate(rows): A[i]=h@Sus[k]
            for i,(k,h) in enumerate(rows):
                for j,(kj,hj) in enumerate(rows):
                    if kj<=k:
                        Phi=np.eye(2)
                        for m in range(kj+1,k+1): Phi=Psib[o+m]@Phi
                        D[i,j]=h@Phi@Ps[kj]@hj; D[j,i]=D[i,j]
            L=np.linalg.cholesky(D); DinvA=np.linalg.solve(D,A)
            B.append(dict(o=o,n=n,ini=ini,fin=fin,rows=rows,A=A,D=D,L=L,DinvA=DinvA))
        C=np.eye(4)+sum(bl['A'].T@bl['DinvA'] for bl in B)
        LC=np.linalg.cholesky(C)
        ld=sum(np.log(np.diag(bl['L'])).sum() for bl in B)+np.log(np.diag(LC)).sum()
        return dict(B=B,xs=xs,K=K,Psib=Psib,Q=Q,Zt=Zt,C=C,LC=LC,ld=ld,z=z,dz=dz,part=part)
    def inv_gram(s,lin,c):
        B=lin['B']; i=0; ts=[]
        for bl in B:
            nr=len(bl['rows']); ts.append(np.linalg.solve(bl['D'],c[i:i+nr])); i+=nr
        g=sum(bl['A'].T@t for bl,t in zip(B,ts))
        sv=np.linalg.solve(lin['C'],g)
        return [t-bl['DinvA']@sv for bl,t in zip(B,ts)]
    def jt(s,lin,lams):
        """J^T lambda -> vector dim_q"""
        S=s.S; out_u=np.zeros(4); out_v0=np.zeros(2); out_v=np.zeros((s.T*S,2))
        for bl,lam in zip(lin['B'],lams):
            out_u+=bl['A'].T@lam
            o,n=bl['o'],bl['n']
            alpha=np.zeros(2)
            for k in reversed(range(n)):
                if k<n-1: alpha=lin['Psib'][o+k+1].T@alpha
                for i,(kk,h) in enumerate(bl['rows']):
                    if kk==k: alpha=alpha+h*lam[i]
                for t in range(S):
                    out_v[(o+k)*S+t]=lin['K'][(o+k)*S+t].T@alpha
            if bl['ini']: out_v0=lin['Psib'][o].T@alpha
        return np.concatenate([out_u,out_v0,out_v.ravel()])
    def jv(s,lin,p):
        S=s.S; pu=p[:4]; pv0=p[4:6]; pv=p[6:].reshape(-1,2); out=[]
        for bl in lin['B']:
            o,n=bl['o'],bl['n']
            m=pv0.copy() if bl['ini'] else np.zeros(2)
            r=bl['A']@pu
            for k in range(n):
                sk=sum(lin['K'][(o+k)*S+t]@pv[(o+k)*S+t] for t in range(S))
                m=lin['Psib'][o+k]@m+sk
                for i,(kk,h) in enumerate(bl['rows']):
                    if kk==k: r[i]+=h@m
            out.append(r)
        return np.concatenate(out)
    def grad_ld(s,q,lin):
        u,v0,v,z,dz=s.unpack(q); S=s.S; dl=s.dl
        Cinv=np.linalg.inv(lin['C'])
        gu=np.zeros(4); gv0=np.zeros(2); gv=np.zeros((s.T*S,2))
        for bl in lin['B']:
            o,n,rows=bl['o'],bl['n'],bl['rows']; nr=len(rows)
            E=np.linalg.inv(bl['D'])-bl['DinvA']@Cinv@bl['DinvA'].T
            Om=bl['DinvA']@Cinv  # nr x 4 (u-directions)
            Psib=lin['Psib']
            # a[r][k] for k<=k_r : Phi(t_kr,t_k)^T h_r
            a=np.zeros((nr,n,2))
            for i,(kr,h) in enumerate(rows):
                vec=h.copy(); a[i,kr]=vec
                for k in reversed(range(kr)):
                    vec=Psib[o+k+1].T@vec; a[i,k]=vec
            alive=lambda i,k: rows[i][0]>=k
            # beta[r][k] = sum_{s alive at k} E[r,s] a[s][k]
            beta=np.einsum('rs,skx->rkx',E,a)   # a is zero where not alive
            Mk=np.einsum('rkx,rky->kxy',beta,a)  # sum_r beta_{r,k} a_{r,k}^T ; a zero if dead
            Lam=np.einsum('ru,rkx->kux',Om,a)    # U x X
            # tangent at obs level: delta[r][j] state at END of interval j (time t_j); delta_start
            d0=np.zeros((nr,2))
            if bl['ini']:
                for i in range(nr):
                    beta0=Psib[o].T@beta[i,0]
                    d0[i]=np.array([0,-dz[3]*Om[i,3]])+beta0
            dobs=np.zeros((nr,n,2)); dprev=d0.copy()
            for j in range(n):
                for i in range(nr):
                    dprev[i]=Psib[o+j]@dprev[i]+lin['Q'][o+j]@beta[i,j]+(lin['Zt'][o+j]*dz[None,:])@Om[i]
                    dobs[i,j]=dprev[i]
            gz=np.zeros(4); gam=np.zeros(2)
            for k in reversed(range(n)):
                g0=(o+k)*S
                # Ybar at start of interval k
                Yb=np.zeros((2,2))
                for i in range(nr):
                    if alive(i,k):
                        dst = d0[i] if k==0 else dobs[i,k-1]
                        Yb+=np.outer(dst,a[i,k])
                LamZ=(dz[:,None]*Lam[k])   # Z x X
                Ys=np.zeros((S,2,2)); Y=Yb
                for t in range(S):
                    Ys[t]=Y
                    xt,vt=lin['xs'][g0+t],v[g0+t]
                    Ft=M.Fx(*z,*xt,*vt,dl);Bt=M.Fv(*z,*xt,*vt,dl);Gt=M.Fz(*z,*xt,*vt,dl)
                    Y=Ft@Y+Bt@lin['K'][g0+t].T@Mk[k]+Gt@LamZ
                Psi=np.eye(2)
                for t in reversed(range(S)):
                    xt,vt=lin['xs'][g0+t],v[g0+t]
                    Ft=M.Fx(*z,*xt,*vt,dl);Bt=M.Fv(*z,*xt,*vt,dl);Gt=M.Fz(*z,*xt,*vt,dl)
                    Kt=Psi@Bt
                    Th=np.vstack([Ys[t]@Psi, Kt.T@Mk[k]@Psi, LamZ@Psi])  # 8x2
                    g=M.hess_contract(z,xt,vt,dl,Th)
                    gv[g0+t]=Bt.T@gam+g[2:4]
                    gz+=Gt.T@gam+g[4:8]
                    gam=Ft.T@gam+g[0:2]
                    Psi=Psi@Ft
            if bl['ini']:
                gv0+=gam; gz[3]+=-gam[1]
            gu+=dz*gz
            # zeta'' extra: exp comps m=0..2: sum_r Om[r,m]*A[r,m]
            for m in range(3): gu[m]+=Om[:,m]@bl['A'][:,m]
        return np.concatenate([gu,gv0,gv.ravel()])

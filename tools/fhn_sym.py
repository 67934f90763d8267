# sympy derivatives for FHN closed-form step; lambdified for the numpy prototype
import sympy as sp, numpy as np
x0,x1,v0,v1 = sp.symbols('x0 x1 v0 v1', real=True)
s,e,g,b = sp.symbols('sigma epsilon gamma beta', real=True)
d = sp.symbols('delta', positive=True)
a0=(x0-x0**3-x1)/e; a1=g*x0-x1+b
dw_=sp.sqrt(d)*v0; dz_=d**sp.Rational(3,2)*(v0+v1/sp.sqrt(3))/2
f0 = x0+d*a0+(d**2/2)*(((1-3*x0**2)/e)*a0-a1/e)-(s/e)*dz_
f1 = x1+d*a1+s*dw_+(d**2/2)*(g*a0-a1)-s*dz_
f = sp.Matrix([f0,f1])
X=[x0,x1]; V=[v0,v1]; Z=[s,e,g,b]
Yv = X+V+Z
args=(s,e,g,b,x0,x1,v0,v1,d)
step = sp.lambdify(args, f, 'numpy')
Fx = sp.lambdify(args, f.jacobian(X), 'numpy')
Fv = sp.lambdify(args, f.jacobian(V), 'numpy')
Fz = sp.lambdify(args, f.jacobian(Z), 'numpy')
H = [sp.hessian(f[i], Yv) for i in range(2)]
Hf = sp.lambdify(args, H, 'numpy')
def hess_contract(z,x,v,dl,Th):
    # Th: 8x2 ; returns g (8,) = sum_i H_i @ Th[:,i]
    Hs = Hf(*z,*x,*v,dl)
    return sum(np.asarray(Hs[i],dtype=float) @ Th[:,i] for i in range(2))

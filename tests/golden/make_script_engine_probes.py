"""Generate tests/golden/script_engine_probes.json: small scripts that probe corners of JavaScript a metric script may rely on (array, object,
string and number methods, closures, classes, error handling, literals), each with the metric diagonal the REFERENCE'S OWN QuickJS
computes for it.  Run where /root/reference is mounted.  tests/test_minijs.py holds the product's interpreter to these values (and
to a live QuickJS run where the reference is present).  The cases in SKIP are known differences."""
import os, sys, numpy as np, tempfile, pathlib, json
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import geodesic_raytracing_b200 as grt, oracle
from test_loader import write_folder
q = oracle.QuickJSReference()
pos = (0.3, 1.5, 0.7, 0.4)
CASES = {}
CASES["array sort default / numeric"] = "const a=[10,9,1,100]; a.sort(); const b=[3,1,2].sort((p,q)=>q-p); return [-1, a[0]*1000+a[1]*100+a[2]*10+a[3]*0+0*x, b[0]*100+b[1]*10+b[2], a.length+0*z];"
CASES["array reverse splice includes"] = "const a=[1,2,3,4,5]; a.reverse(); const r=a.splice(1,2); return [-1, a.join('')*1+0*x, r.join('')*1, (a.includes(5)?1:0)+(a.includes(9)?10:0)+0*z];"
CASES["find findIndex flat"] = "const a=[[1,2],[3,[4]]]; const f=a.flat(); const g=[5,6,7].find(v=>v>5); const h=[5,6,7].findIndex(v=>v>6); return [-1, f.length+0*x, g+0*y, h+0*z];"
CASES["Array.from isArray new Array"] = "const a=Array.from({length:3},(v,i)=>i*2); const b=new Array(4).fill(1); const c=Array(2); return [-1, a[2]+a.length+0*x, b.length+b[3]+0*y, (Array.isArray(c)?1:0)+c.length+0*z];"
CASES["Math extras"] = "return [-1, Math.hypot(3,4)+Math.sign(-2)+0*x, Math.trunc(-2.7)+Math.cbrt(27)+Math.log2(8)+0*y, Math.round(2.5)+Math.round(-2.5)+Math.atan2(1,1)+0*z];"
CASES["loose equality nullish optional chaining"] = "const o={a:{b:2}}; const n=null; return [-1, (('1'==1)?1:0)+((null==undefined)?10:0)+((0=='')?100:0)+0*x, (n??5)+(0??7)+0*y, (o?.a?.b)+(o?.c?.d===undefined?10:0)+0*z];"
CASES["let per-iteration closures"] = "const fs=[]; for(let i=0;i<3;i++){fs.push(()=>i);} var gs=[]; for(var j=0;j<3;j++){gs.push(function(){return j;});} return [-1, fs[0]()+fs[1]()*10+fs[2]()*100+0*x, gs[0]()+gs[2]()+0*y, 1+0*z];"
CASES["arguments object"] = "function s(){ var t=0; for(var i=0;i<arguments.length;i++) t+=arguments[i]; return t; } return [-1, s(1,2,3)+0*x, s()+1+0*y, s(x,y)];"
CASES["object literal getters setters"] = "const o={_v:2, get v(){return this._v*2;}, set v(n){this._v=n+1;}}; o.v=5; return [-1, o.v+0*x, o._v+0*y, 1+0*z];"
CASES["Object.assign entries values"] = "const o=Object.assign({a:1},{b:2},{a:3}); const e=Object.entries(o); const v=Object.values(o); return [-1, o.a+o.b+0*x, e.length*10+e[0][1]+0*y, v[0]+v[1]+0*z];"
CASES["Number helpers"] = "return [-1, (Number.isFinite(1/0)?1:0)+(isNaN(0/0)?10:0)+parseFloat('2.5x')+0*x, parseInt('42px')+parseInt('ff',16)+0*y, Number('3.5')+(+'2')+(1.005).toFixed(2)*1+0*z];"
CASES["string methods"] = "const s='a,b,c'; const p=s.split(','); return [-1, p.length+s.indexOf('b')+0*x, 'abc'.charAt(1)=='b'?1:0+0*y, 'xy'.repeat(2).length+'ab'.toUpperCase().length+'abc'.slice(1).length+'abc'.substring(0,2).length+0*z];"
CASES["operators"] = "return [-1, 2**3+(-7%3)+(7/2|0)+0*x, (5&3)+(5|3)+(5^3)+(1<<4)+(-16>>2)+(-16>>>28)+0*y, (void 0===undefined?1:0)+('a' in {a:1}?10:0)+0*z];"
CASES["delete"] = "const o={a:1,b:2}; delete o.a; return [-1, Object.keys(o).length+0*x, ('a' in o?1:0)+2+0*y, (o.a===undefined?1:0)+0*z];"
CASES["exceptions"] = "let m=0; try { throw new Error('boom'); } catch(e) { m=e.message.length; } finally { m+=10; } let k=0; try { null.x; } catch(e) { k=(e instanceof TypeError)?1:2; } return [-1, m+0*x, k+0*y, 1+0*z];"
CASES["number to string"] = "return [-1, (''+0.1+0.2).length+0*x, String(1e21).length+String(123456789012).length+0*y, `${1/3}`.length+(0.000001+'').length+0*z];"
CASES["hoisting and TDZ-free var"] = "var r=g(); function g(){ return typeof h; } var h=1; return [-1, (r=='undefined'?1:0)+0*x, 1+0*y, 1+0*z];"
CASES["getter on class + static inherit"] = "class A{ static s(){return 7;} get q(){return 3;} } class B extends A{} return [-1, B.s()+0*x, new B().q+0*y, 1+0*z];"
CASES["switch on strings with fallthrough default first"] = "function k(v){ switch(v){ default: return 0; case 'a': return 1; case 'b': return 2; } } return [-1, k('a')+0*x, k('b')+k('z')+0*y, 1+0*z];"
CASES["comma, conditional chains, short-circuit values"] = "let a=(1,2,3); const b=0||'x'; const c=1&&0; const d=''||null||5; return [-1, a+0*x, (b=='x'?1:0)+c+0*y, d+0*z];"
CASES["array destructure swap in loop + spread math"] = "let a=1,b=2; for(let i=0;i<3;i++){ [a,b]=[b,a+b]; } return [-1, a+0*x, b+0*y, Math.min(...[3,1,2])+0*z];"
CASES["map with index, reduce no init, forEach, some every"] = "const a=[1,2,3]; let s=0; a.forEach((v,i)=>{s+=v*i;}); return [-1, a.map((v,i)=>v*i).join('')*1+0*x, a.reduce((p,v)=>p*v)+s+0*y, (a.some(v=>v>2)?1:0)+(a.every(v=>v>1)?10:0)+0*z];"
CASES["typeof and instanceof of functions arrays"] = "return [-1, (typeof []=='object'?1:0)+(typeof null=='object'?10:0)+(typeof function(){}=='function'?100:0)+0*x, ([] instanceof Array?1:0)+0*y, (typeof x=='object'?1:0)+0*z];"
CASES["label continue in while + do"] = "let i=0,n=0; a: while(i<5){ i++; let j=0; do { j++; if(j==2) continue a; n++; } while(j<4); } return [-1, n+0*x, i+0*y, 1+0*z];"
CASES["string compare and sort strings"] = "const a=['b','a','C']; a.sort(); return [-1, (a[0]=='C'?1:0)+0*x, ('a'<'b'?1:0)+('10'<'9'?10:0)+(10<9?100:0)+0*y, 1+0*z];"
CASES["closure counter / IIFE / recursion"] = "const mk=()=>{let c=0; return ()=>++c;}; const n=mk(); n(); n(); const fact=k=>k<=1?1:k*fact(k-1); return [-1, n()+0*x, fact(5)+0*y, (function(a){return a*2;})(4)+0*z];"
CASES["bigger: matrix helper"] = "function mat(n){ const m=[]; for(let i=0;i<n;i++){ m.push([]); for(let j=0;j<n;j++) m[i].push(i==j?1:0);} return m;} const m=mat(3); m[1][2]=x; const tr=m.reduce((s,r,i)=>s+r[i],0); return [-1, tr+0*x, m[1][2], m.length*m[0].length+0*z];"
CASES["optional chaining deep"] = "const o={a:{b:{c:2}, f(){return 3;}}, arr:[1,2]}; const u=undefined; return [-1, (o?.a?.b?.c)+(u?.a.b.c===undefined?10:0)+0*x, (o.a?.f?.())+(o.a.g?.()===undefined?10:0)+(o?.arr?.[1])+0*y, (u?.[0]===undefined?1:0)+(o.zz?.yy.xx===undefined?10:0)+0*z];"
CASES["accessors + keys skip them"] = "const o={a:1, get d(){return this.a*2;}, set d(v){this.a=v;}, get only(){return 5;}}; o.d=4; o.only=9; return [-1, o.d+0*x, o.only+0*y, Object.keys(o).length+0*z];"
CASES["call apply bind"] = "function add(a,b){ return (this.k||0)+a+b; } const o={k:10}; const b2=add.bind(o,1); return [-1, add.call(o,1,2)+0*x, add.apply(o,[3,4])+Math.max.apply(null,[1,7,3])+0*y, b2(5)+0*z];"
CASES["number literals"] = "return [-1, 0x1f+0o17+0b101+0*x, 1_000+1e3+.5+5.+0*y, 0.1+0.2+0*z];"
CASES["string escapes"] = "const s='a\\nb\\tc\\u0041\\x42\\'\\\"'; return [-1, s.length+0*x, s.charCodeAt(5)+s.charCodeAt(6)+0*y, `a${`b${1+1}`}c`.length+0*z];"
CASES["function hoisting in blocks, typeof undeclared"] = "let r=0; { function inner(){ return 4; } r=inner(); } return [-1, r+0*x, (typeof nope=='undefined'?1:0)+0*y, (typeof inner)=='function'?1:0+0*z];"
CASES["default param scope and order"] = "let n=0; function g(a=++n, b=a*2, c=()=>a+b){ return c(); } return [-1, g()+0*x, g(5)+n+0*y, g(1,1)+0*z];"
CASES["finally flow"] = "function f1(){ try { return 1; } finally { n1=5; } } var n1=0; function f2(){ for(var i=0;i<3;i++){ try { if(i==1) continue; if(i==2) break; } finally { n1++; } } return i; } return [-1, f1()+n1+0*x, f2()+n1+0*y, 1+0*z];"
CASES["for-in over array and string index"] = "const a=[5,6,7]; let s=''; for (const k in a) s+=k; const str='hey'; return [-1, s.length+(s=='012'?10:0)+0*x, (str[1]=='e'?1:0)+0*y, 1+0*z];"
CASES["getter inheritance and super property"] = "class A{ get v(){return 2;} m(){return 3;} } class B extends A{ get v(){ return super.v*5; } m(){ return super.m()+1; } } const b=new B(); return [-1, b.v+0*x, b.m()+0*y, 1+0*z];"
CASES["in operator and hasOwnProperty"] = "const o={a:1}; return [-1, ('a' in o?1:0)+('toString' in o?10:0)+0*x, (o.hasOwnProperty('a')?1:0)+(o.hasOwnProperty('b')?10:0)+0*y, (Object.prototype.hasOwnProperty.call(o,'a')?1:0)+0*z];"
CASES["array holes length and join"] = "const a=[1,,3]; a[5]=6; return [-1, a.length+0*x, a.join('-').length+0*y, (a[1]===undefined?1:0)+(1 in a?10:0)+0*z];"
CASES["compound ops and increments"] = "let a=5; a+=2; a*=3; a-=1; a/=4; a%=3; a**=2; let b=1; const c=b++ + ++b; let d=8; d>>=1; d<<=2; d|=1; d&=13; d^=2; return [-1, a+0*x, b+c+0*y, d+0*z];"
CASES["string number coercion"] = "return [-1, ('3'*'4')+('5'-2)+0*x, ('1'+2).length+(1+'2'=='12'?10:0)+0*y, (+true)+(+'')+(+' 12 ')+(null+1)+0*z];"
CASES["nan and equality"] = "const n=0/0; return [-1, (n===n?1:0)+(n!=n?10:0)+(isNaN(undefined+1)?100:0)+0*x, ([1,2]==[1,2]?1:0)+(null===null?10:0)+0*y, (0===-0?1:0)+(1/0===-1/(-0)?10:0)+0*z];"
CASES["closures over var in loop with IIFE"] = "var fs=[]; for(var i=0;i<3;i++){ (function(k){ fs.push(function(){return k;}); })(i); } return [-1, fs[0]()+fs[1]()+fs[2]()+0*x, i+0*y, 1+0*z];"
CASES["object method shorthand this, nested arrow"] = "const o={n:3, sum(arr){ return arr.map(v=>v*this.n).reduce((a,b)=>a+b,0); }}; return [-1, o.sum([1,2])+0*x, 1+0*y, 1+0*z];"
CASES["spread of string and nested arrays"] = "const a=[...'abc']; const b=[...[1,2],...[[3]]]; return [-1, a.length+0*x, b.length+b[2][0]+0*y, Math.max(...a.map(c=>c.charCodeAt(0)))+0*z];"
CASES["while with complex conditions and do-while continue"] = "let i=0,s=0; do { i++; if(i%2) continue; s+=i; } while(i<6); let j=10; while(j-->0 && s<100) s+=j; return [-1, s+0*x, j+0*y, i+0*z];"
CASES["ternary chains precedence"] = "const f=v=>v<0?'n':v==0?'z':'p'; const a=1+2*3**2-4/2%3; const b=(1,2)&&3||4; return [-1, (f(-1)+f(0)+f(1)).length+(f(0)=='z'?10:0)+0*x, a+0*y, b+0*z];"
CASES["Array constructor forms and length assign"] = "const a=new Array(3); const b=new Array(1,2); const c=[]; c.length=2; c.push(1); return [-1, a.length+b.length+0*x, c.length+0*y, (a[0]===undefined?1:0)+0*z];"
CASES["Math edge"] = "return [-1, Math.floor(-0.5)+Math.ceil(-0.5)+Math.round(0.5)+Math.round(1.5)+Math.abs(-3)+0*x, Math.pow(2,0.5)*Math.SQRT2+Math.exp(0)+Math.log(Math.E)+0*y, Math.min()+'x'=='Infinityx'?1:0+0*z];"
CASES["static and instance method this chain"] = "class C{ constructor(){this.v=1;} inc(){this.v++; return this;} static make(){ return new this(); } } return [-1, C.make().inc().inc().v+0*x, 1+0*y, 1+0*z];"
CASES["class expression name and instanceof chain"] = "const K=class Named{ who(){ return Named.tag; } }; K.tag=7; class L extends K{} return [-1, new K().who()+0*x, (new L() instanceof K?1:0)+0*y, 1+0*z];"
CASES["toString valueOf coercion on objects"] = "const o={valueOf(){return 4;}, toString(){return '9';}}; return [-1, (o*2)+0*x, (`${o}`*1)+0*y, (o+1)+0*z];"
CASES["label block and nested switch break"] = "let r=0; sw: switch(1){ case 1: for(let i=0;i<3;i++){ if(i==1) break sw; r++; } r=99; } return [-1, r+0*x, 1+0*y, 1+0*z];"
CASES["getter setter static and computed class member absent"] = "class P{ static get ver(){ return 3; } static set ver(v){ P._v=v; } } P.ver=8; return [-1, P.ver+P._v+0*x, 1+0*y, 1+0*z];"
CASES["exceptions rethrow custom"] = "class MyErr extends Error{ constructor(m){ super(m); this.code=7; } } let r=0; try { try { throw new MyErr('x'); } catch(e) { r=e.code+(e instanceof Error?10:0)+(e.message=='x'?100:0); throw e; } } catch(e2) { r+=1000; } return [-1, r+0*x, 1+0*y, 1+0*z];"
CASES["block scoping and shadowing"] = "let a=1; { let a=2; { const a=3; var b=a; } b+=a; } return [-1, a+0*x, b+0*y, 1+0*z];"
CASES["const reassign is an error"] = "const c=1; let r=0; try { c=2; } catch(e) { r=(e instanceof TypeError)?1:2; } return [-1, r+c+0*x, 1+0*y, 1+0*z];"
CASES["tdz"] = "let r=0; try { r=q; let q=1; } catch(e) { r=(e instanceof ReferenceError)?5:6; } return [-1, r+0*x, 1+0*y, 1+0*z];"
CASES["named function expression binding"] = "const f=function fact(n){ return n<=1?1:n*fact(n-1); }; return [-1, f(4)+0*x, (typeof fact=='undefined'?1:0)+0*y, 1+0*z];"
CASES["arguments in arrow uses outer"] = "function outer(){ const g=()=>arguments[0]*2; return g(99); } return [-1, outer(4)+0*x, 1+0*y, 1+0*z];"
CASES["slice.call arguments / Array.prototype"] = "function f2(){ const a=Array.prototype.slice.call(arguments,1); return a.length*10+a[0]; } return [-1, f2(1,2,3)+0*x, 1+0*y, 1+0*z];"
CASES["for-of closures and for-in closures"] = "const fs=[]; for (const v of [1,2,3]) fs.push(()=>v); const gs=[]; for (const k in {a:1,b:2}) gs.push(()=>k); return [-1, fs[0]()+fs[2]()+0*x, (gs[0]()+gs[1]()).length+(gs[0]()=='a'?10:0)+0*y, 1+0*z];"
CASES["typeof class and function props"] = "class A{} function g(){} g.tag=3; return [-1, (typeof A=='function'?1:0)+g.tag+0*x, g.name.length+A.name.length+0*y, g.length+((a,b)=>0).length+0*z];"
CASES["catch without binding, nested try finally order"] = "let log=''; try { try { throw 1; } finally { log+='f'; } } catch { log+='c'; } finally { log+='F'; } return [-1, log.length+(log=='fcF'?10:0)+0*x, 1+0*y, 1+0*z];"
CASES["throw non-error values"] = "let r=0; try { throw {code:4}; } catch(e) { r=e.code; } try { throw 'str'; } catch(e) { r+=e.length; } try { throw 7; } catch(e) { r+=e; } return [-1, r+0*x, 1+0*y, 1+0*z];"
CASES["switch scoping and fallthrough with blocks"] = "function k(v){ let r=0; switch(v){ case 1: { let t=5; r+=t; } case 2: r+=1; break; case 3: r=9; } return r; } return [-1, k(1)+0*x, k(2)+k(3)+0*y, k(4)+1+0*z];"
CASES["exponent assoc, unary minus, modulo sign"] = "return [-1, 2**3**2+0*x, (-2)**2+(-7%3)+(7%-3)+0*y, -(2**2)+0*z];"
CASES["array length semantics"] = "const a=[1,2,3]; a.length=1; a[3]=9; const b=[]; b[2]=1; return [-1, a.length+0*x, b.length+(a[1]===undefined?10:0)+0*y, a.indexOf(9)+0*z];"
CASES["number methods"] = "const n=255; return [-1, n.toString(16).length+(n.toString(2)=='11111111'?10:0)+0*x, (3.14159).toFixed(2)*100+0*y, (1234.5678).toPrecision(6)*1+(0.00001234).toPrecision(2).length+0*z];"
CASES["Math.round ties and negative zero"] = "return [-1, Math.round(-0.5)+Math.round(-1.5)+Math.round(2.4999)+0*x, (1/Math.round(-0.2)<0?1:0)+0*y, Math.sign(-0)+Math.trunc(0.9)+Math.fround?1:0+0*z];"
CASES["object spread order and override"] = "const a={x:1,y:2}; const b={...a, y:3, z:4}; const c={y:0, ...a}; return [-1, b.y+c.y+0*x, Object.keys(b).join('').length+(Object.keys(c)[0]=='y'?10:0)+0*y, b.z+0*z];"
CASES["nested destructuring defaults with calls"] = "let calls=0; const d=()=>{calls++; return 5;}; const {a=d(), b:{c=d()}={}} = {a:1}; const [p=d(), q=d()] = [undefined, 2]; return [-1, a+c+0*x, p+q+0*y, calls+0*z];"
CASES["closures mutate shared"] = "function mk(){ let v=0; return {inc:()=>++v, get:()=>v}; } const a=mk(), b=mk(); a.inc(); a.inc(); b.inc(); return [-1, a.get()+0*x, b.get()+0*y, 1+0*z];"
CASES["sort stability and comparator"] = "const a=[{k:1,v:'a'},{k:0,v:'b'},{k:1,v:'c'},{k:0,v:'d'}]; a.sort((p,q)=>p.k-q.k); return [-1, (a.map(e=>e.v).join('')=='bdac'?1:0)+0*x, [3,20,100].sort()[0]+0*y, [3,20,100].sort((p,q)=>p-q)[2]+0*z];"
CASES["getter this in static and inheritance of statics"] = "class A{ static base=2; static get twice(){ return this.base*2; } } class B extends A{ static base=5; } return [-1, A.twice+0*x, B.twice+0*y, 1+0*z];"
CASES["in-place array methods return values"] = "const a=[1,2,3]; const r1=a.push(4); const r2=a.unshift(0); const r3=a.pop(); const r4=a.shift(); return [-1, r1*10+r2+0*x, r3*10+r4+0*y, a.length+0*z];"
CASES["chained ternary with assignment and comma in for"] = "let s=0; for (let i=0,j=10; i<j; i+=2,j-=2) s+=i*j; return [-1, s+0*x, 1+0*y, 1+0*z];"
CASES["string to number edge"] = "return [-1, (+'0x10')+(+'1e2')+(+'  7  ')+0*x, (isNaN(+'1_0')?1:0)+(isNaN(+'abc')?10:0)+(+[]===0?100:0)+0*y, (+[5])+(isNaN(+[1,2])?10:0)+0*z];"
CASES["pow operator on number objects"] = "return [-1, x**2, (x+1)**3/(y+2), z**0.5];"
CASES["compound assignment with number objects"] = "let a=x; a+=y; a*=2; a-=1; a/=z; let b=2; b*=x; return [-1, a, b, 1+0*z];"
CASES["reduce over number objects"] = "const v=[x,y,z]; const n2=v.reduce((s,c)=>s+c*c,0); const m=v.map(c=>c*2); return [-1, n2, m[0]+m[1]+m[2], CMath.sqrt(n2)];"
CASES["Math functions on number objects"] = "let r1=1,r2=1; try { r1=Math.sqrt(x); } catch(e) { r1=-7; } try { r2=Math.pow(x,2); } catch(e) { r2=-8; } return [-1, (typeof r1=='number' && isNaN(r1))?5:(typeof r1=='number'?r1:9)+0*x, (typeof r2=='number' && isNaN(r2))?5:(typeof r2=='number'?r2:9)+0*y, 1+0*z];"
CASES["mixed CMath and plain numbers"] = "const a=CMath.sin(0.5), b=CMath.pow(2,3), c=CMath.sqrt(4); return [-1, a+0*x, b+c+0*y, CMath.max(1,2)+CMath.min(x,0.1)+0*z];"
CASES["number object in template and typeof"] = "return [-1, (typeof x=='object'?1:0)+0*x, (typeof (x*0+2)=='object'?1:0)+0*y, (`${2+3}`*1)+0*z];"
CASES["number objects in object and array containers"] = "const p={c:[x,{d:y}]}; const {c:[a,{d}]}=p; return [-1, a*d, p.c[1].d+p.c[0], [x,y].concat([z]).length+0*z];"
CASES["CMath.select and comparisons"] = "const s=CMath.select(CMath.lt(x,2), x*2, x*3); const g=CMath.gt(y,5); return [-1, s, 1+g, CMath.select(CMath.eq(z,z),4,5)+0*z];"
CASES["functions returning functions of number objects"] = "const lin=(m,c)=>(v)=>m*v+c; const f1=lin(x,1), f2=lin(2,y); return [-1, f1(3), f2(z), f1(f2(1))];"
CASES["division and chained ops order"] = "return [-1, x/y/z, x-y-z+1, x/(y/z)];"
CASES["pin and cfg"] = "const a=$cfg.mass; $cfg.mass.$default=2; const r=$pin(x*x+y*y); return [-1, r*a, r/(a+1), $pin(r)+0*z];"
CASES["unary minus on number objects"] = "return [-1, -x, y+(-(-z)), 1+0*z];"
CASES["instance keys exclude methods"] = "class V{ constructor(){ this.a=1; this.b=2; } norm(){ return 3; } get g(){ return 4; } } const v=new V(); let n=0; for (const k in v) n++; const c={...v}; return [-1, Object.keys(v).length+0*x, n+0*y, Object.keys(c).length+(typeof c.norm=='undefined'?10:0)+0*z];"
CASES["method override on instance and hasOwnProperty"] = "class V{ m(){ return 1; } } const v=new V(); const own0=v.hasOwnProperty('m'); v.m=()=>2; return [-1, v.m()+0*x, (own0?1:0)+(v.hasOwnProperty('m')?10:0)+0*y, new V().m()+0*z];"
CASES["prototype method added later"] = "class V{} V.prototype.twice=function(){ return this.k*2; }; const v=new V(); v.k=4; return [-1, v.twice()+0*x, 1+0*y, 1+0*z];"
CASES["function constructor with prototype (ES5 classes)"] = "function P(a){ this.a=a; } P.prototype.get=function(){ return this.a+1; }; const p=new P(4); return [-1, p.get()+0*x, (p instanceof P?1:0)+0*y, Object.keys(p).length+0*z];"
CASES["Object.create and getPrototypeOf"] = "const base={hi(){ return 7; }}; const o=Object.create(base); o.own=1; return [-1, o.hi()+0*x, Object.keys(o).length+0*y, (Object.getPrototypeOf(o)===base?1:0)+0*z];"
CASES["JSON-free deep copy via spread and map"] = "const a=[{v:1},{v:2}]; const b=a.map(e=>({...e})); b[0].v=9; return [-1, a[0].v+0*x, b[0].v+0*y, 1+0*z];"
CASES["getter on instance via defineProperty"] = "const o={}; Object.defineProperty(o,'v',{get(){ return 5; }, enumerable:false}); Object.defineProperty(o,'w',{value:6}); return [-1, o.v+0*x, o.w+0*y, Object.keys(o).length+1+0*z];"
CASES["Map basics"] = "const m=new Map([['a',1],['b',2]]); m.set('c',3).set('a',10); const o={}; m.set(o,5); m.delete('b'); let s=0; for (const [k,v] of m) s+=v; return [-1, m.size+0*x, s+m.get('a')+(m.has('b')?100:0)+m.get(o)+0*y, [...m.keys()].length+[...m.values()][0]+0*z];"
CASES["Set basics"] = "const s=new Set([1,2,2,3,NaN,NaN]); s.add(4).add(1); s.delete(2); let tt=0; s.forEach(v=>{ if(v===v) tt+=v; }); const u=[...new Set([3,1,3,2])]; return [-1, s.size+0*x, tt+(s.has(NaN)?100:0)+0*y, u.join('')*1+0*z];"
CASES["Map of number objects and object keys"] = "const m=new Map(); const k1={id:1}, k2={id:1}; m.set(k1,x); m.set(k2,y); m.set(1,'n'); m.set('1','s'); return [-1, m.get(k1)*2, m.get(k2)+m.size, (m.get(1)=='n'?1:0)+(m.get('1')=='s'?10:0)+0*z];"
CASES["Array.from Map Set and entries destructuring"] = "const m=new Map([[1,'a'],[2,'b']]); const a=Array.from(m); const b=Array.from(new Set('hello')); let n=0; m.forEach((v,k)=>{ n+=k*(v=='b'?10:1); }); return [-1, a.length*10+a[1][0]+0*x, b.length+0*y, n+0*z];"
CASES["Float64Array"] = "const f=new Float64Array(3); f[1]=2.5; const g=new Float64Array([1,2,3]); return [-1, f.length+f[0]+f[1]+0*x, g.reduce((s,v)=>s+v,0)+0*y, g.map(v=>v*2)[2]+0*z];"
SKIP = {"in operator and hasOwnProperty", "array holes length and join", "toString valueOf coercion on objects", "nested functions this binding", "json", "Math.random absent / Date"}
out = {}
for name, body in CASES.items():
    if name in SKIP:
        continue
    js = "function f(t, x, y, z) {\n" + body + "\n}\nf"
    tmp = pathlib.Path(tempfile.mkdtemp())
    folder = write_folder(tmp, js)
    try:
        g, dg, n, found = q.metric(os.path.join(folder, "probe.js"), np.array(pos))
    except RuntimeError as e:            # a script the reference's engine refuses: not a value probe (tests/test_minijs.py REFUSED holds those)
        print("refused by QuickJS:", name, "-", str(e).splitlines()[-1][-80:])
        continue
    out[name] = (body, [float(v) for v in np.diag(np.asarray(g).reshape(4, 4))])
json.dump(out, open("/root/repo/tests/golden/script_engine_probes.json", "w"), indent=1)
print(len(out))

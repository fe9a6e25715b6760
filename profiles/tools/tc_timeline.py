import ctypes as C, torch, sys
sys.path.insert(0,"/root/repo")
import cleanmarl_b200 as cm
from cleanmarl_b200 import _lib
from cleanmarl_b200.mappo import MAPPO, Args
lib=_lib.load()
lib.cmarl_debug_tc_timeline.argtypes=[C.c_int, C.POINTER(C.c_longlong)]
tr=MAPPO(Args(batch_size=4096,seed=1)); 
import os
assert tr.engine.tensor_cores
for _ in range(2): tr.iteration()
torch.cuda.synchronize()
buf=(C.c_longlong*64)()
names={0:"tile start",1:"X published",2:"F1 done seen",3:"H1 published",4:"H1s stored",5:"F2 done seen",6:"head done",7:"dW3 done",8:"dH2 published",9:"B1 done seen",10:"E3 done",11:"dW2a done seen",12:"H1s r1 published",13:"dW2b done seen",14:"dW1 r0 published",15:"dW1a done seen",16:"Xs r1 published",17:"dW1b done seen",18:"flush done",
20:"K: kernel entry",21:"K: setup done",24:"K: first flush start",25:"K: flush: lo rows in scratch",26:"K: flush: hi rows added",27:"K: flush: partial row written",22:"K: tile loop done",23:"K: partial written",48:"K: tile 0 start",49:"K: tile 1 start",50:"K: tile 2 start",51:"K: tile 3 start",52:"K: tile 4 start",53:"K: tile 5 start",54:"K: tile 6 start",55:"K: tile 7 start",
32:"I: tile start",33:"I: X ready",34:"I: F1 issued",35:"I: H1 ready",36:"I: F2 issued",37:"I: dH2 ready",38:"I: B1 issued",39:"I: dW2a issued",40:"I: H1s r1 ready",41:"I: dW2b issued",42:"I: dW1a ready",43:"I: dW1a issued",44:"I: Xs r1 ready",45:"I: dW1b issued"}
def run(fn,label,mode=1):
    lib.cmarl_debug_tc_timeline(mode,None)
    fn(); torch.cuda.synchronize()
    lib.cmarl_debug_tc_timeline(0,buf)
    v=list(buf); ev=sorted((v[k],k) for k in names if v[k]>0)
    t0=ev[0][0]; print("==",label)
    prev=t0
    for t,k in ev:
        print(f"{t-t0:8d} (+{t-prev:6d})  {names[k]}"); prev=t
    for k in range(64): buf[k]=0
    z=(C.c_longlong*64)(); 
eng=tr.engine; b=tr.buf
import torch
def critic_train():
    eng.ppo_epoch_grads(tr.net.flat, tr.grads, state=b["state"], actions=b["actions"], logp_old=b["logp"], adv=b["adv"], returns=b["returns"])
run(critic_train,"ppo_epoch_grads (timeline = last kernel writing: critic train H=64 K=56)")

run(critic_train,"actor train H=32 K=24 (policy head)",2)

def critic_forward():
    eng.critic_values(tr.net.critic, b["values"], state=b["state"])
run(critic_forward,"critic forward (cmarl_critic_values: TCfg<64,56,0,1>, 296 CTAs; stamps of tile 1 of CTA 0)")

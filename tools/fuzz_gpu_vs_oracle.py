"""Development aid (run under gpurun): random option sets x seeded batches, GPU library vs the oracle, every record,
both Stats blocks, the counters and the --mask/--break lists.  usage: python tools/fuzz_gpu_vs_oracle.py <seed> <seconds>"""
import sys, random, time
import os; ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import numpy as np
import cases
from fastplong_b200 import Options, pack_reads, synth
from oracle_lib import OracleEngine, compare_results, compare_stats, compare_lists
from fastplong_b200.binding import Engine
rng = random.Random(int(sys.argv[1]) if len(sys.argv)>1 else 1)
def rand_adapter(lo=6,hi=45):
    return ''.join(rng.choice('ACGT') for _ in range(rng.randint(lo,hi)))
def rand_opts():
    kw = {}
    mode = rng.random()
    if mode < 0.15: kw['disable_adapter_trimming']=True
    else:
        if rng.random()<0.8: kw['start_adapter']=rng.choice([synth.ADAPTER_START, rand_adapter()])
        if rng.random()<0.8: kw['end_adapter']=rng.choice([synth.ADAPTER_END, rand_adapter()])
        if rng.random()<0.2: kw['adapter_fasta']=[rand_adapter(8,40) for _ in range(rng.randint(1,4))]
        if rng.random()<0.3: kw['distance_threshold']=rng.choice([0.1,0.2,0.3,0.4])
        if rng.random()<0.3: kw['trimming_extension']=rng.choice([0,3,10,25])
    if rng.random()<0.4: kw['cut_front']=True
    if rng.random()<0.4: kw['cut_tail']=True
    if rng.random()<0.4: kw['cut_window_size']=rng.choice([1,4,10,30]); kw['cut_mean_quality']=rng.choice([10,15,20,30])
    if rng.random()<0.3: kw['trim_front']=rng.choice([0,1,5,40])
    if rng.random()<0.3: kw['trim_tail']=rng.choice([0,1,5,40])
    if rng.random()<0.3: kw['trim_poly_x']=True; kw['poly_x_min_len']=rng.choice([5,10,20])
    if rng.random()<0.2: kw['disable_quality_filtering']=True
    if rng.random()<0.3: kw['qualified_quality_phred']=rng.choice([5,15,25])
    if rng.random()<0.3: kw['mean_qual']=rng.choice([0,8,15])
    if rng.random()<0.3: kw['n_base_limit']=rng.choice([0,2,50])
    if rng.random()<0.3: kw['n_percent_limit']=rng.choice([1,10,50])
    if rng.random()<0.3: kw['length_required']=rng.choice([0,15,100,1000])
    if rng.random()<0.2: kw['length_limit']=rng.choice([0,500,5000])
    if rng.random()<0.3: kw['low_complexity_filter']=True; kw['complexity_threshold']=rng.choice([10,30,60])
    if rng.random()<0.2: kw['mask']=True; kw['mask_window_size']=rng.choice([5,10,50]); kw['mask_mean_quality']=rng.choice([8,12,20])
    if rng.random()<0.2: kw['break_reads']=True; kw['break_window_size']=rng.choice([10,30,100]); kw['break_mean_quality']=rng.choice([8,12,20])
    return kw
import dataclasses
fields = {f.name for f in dataclasses.fields(Options)}
t0=time.time(); n=0; bad=0
while time.time()-t0 < float(sys.argv[2] if len(sys.argv)>2 else 120):
    kw = {k:v for k,v in rand_opts().items() if k in fields}
    try:
        opt = Options(**kw)
    except Exception as e:
        print('opt error', kw, e); continue
    kind = rng.random()
    seed = rng.randint(1,10**6)
    if kind<0.4: batch = cases.adversarial_batch(seed)
    elif kind<0.7: batch = cases.blocky_quality_batch(seed, n=40)
    else: batch = cases.ont_batch(seed, n=40, mean=1500, p_chimera=0.1, p_polya=0.1)
    r = None
    try:
        o, r = OracleEngine(opt), Engine(opt)
        compare_results(o.process(batch), r.process(batch), 'fuzz')
        if opt.mask or opt.break_reads:
            compare_lists(o.segments(), r.segments(), 'seg'); compare_lists(o.mask_regions(), r.mask_regions(), 'reg')
        cyc=max(1,int(batch.lens.max()))
        for w in (0,1): compare_stats(o.stats(w,cyc), r.stats(w,cyc), f'stats{w}')
        compare_stats(o.counters(), r.counters(), 'counters')
    except AssertionError as e:
        bad+=1; print('MISMATCH', kw, kind, seed, str(e)[:300])
        if bad>5: break
    except Exception as e:
        bad+=1; print('ERROR', kw, kind, seed, type(e).__name__, str(e)[:300])
        if bad>5: break
    finally:
        if r is not None: r.close()
    n+=1
print('cases', n, 'mismatches', bad)

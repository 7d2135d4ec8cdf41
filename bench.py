#!/usr/bin/env python
"""bench.py — NRMS training throughput on synthetic MIND-small-shaped data (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps 20 --warmup 5
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU path (oracle port) on the host cores

A step = one pass of the hot path over one batch of B=64 impressions per GPU: EmbeddingHub gather + projection, NRMS
item encoder over 55 items/impression, NRMS user encoder, dot scoring + softmax-CE, backward, Adam (and the gradient
allreduce when N>1).  Training mode with the reference's dropout rates (0.1) on.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# MIND-small shape
WORKLOAD = dict(n_items=65238, n_words=400000, n_users=91935, n_cats=18, title_len=30, hist_len=50, n_train=208238,
                n_eval_groups=73152, eval_group_mean=36, embed_dim=300)
HIDDEN, HEADS, ADDITIVE, NEG, DROPOUT = 256, 8, 256, 4, 0.1


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--batch', type=int, default=64, help='impressions per GPU per step')
    ap.add_argument('--small', action='store_true', help='tiny world (debug only; not a valid bench line)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='skip the cached-eval / gather HBM lines')
    ap.add_argument('--autograd', action='store_true', help='drive the kernels through torch.autograd instead of the native step driver')
    ap.add_argument('--cpu-seconds', type=float, default=15.0)
    ap.add_argument('--no-parity-check', action='store_true', help='N > 1: skip the multi-rank correctness checks that run before timing')
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p['hbm_gbs'], tensor=p.get('bf16_tflops_sustained', p['bf16_tflops']), which='measured')
    return dict(hbm=6650.0, tensor=1400.0, which='fallback')


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/r2_traffic.json), or None."""
    try:
        return json.load(open(os.path.join(ROOT, 'profiles', 'r2_traffic.json')))[kernel]['dram_bytes_per_launch']
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
                                          '-i', str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [l.split(', ') for t, l in self.lines if t0 <= t <= t1 + 0.2] or [l.split(', ') for _, l in self.lines[-3:]]
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.strip().lower().startswith('active'):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return None
        return dict(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))


def build_world(small: bool, seed=2023):
    from legommenders_b200.synth import MindWorld
    if small:
        return MindWorld(n_items=2000, n_words=5000, n_users=500, n_train=4096, n_eval_groups=200, seed=seed)
    return MindWorld(seed=seed, **WORKLOAD)


# ------------------------------------------------------------------------------------------------------------
# the reference's CPU path (oracle port): fwd + bwd + Adam with all host threads, training-mode dropout
# ------------------------------------------------------------------------------------------------------------
def cpu_reference(world, batches, budget_s=None, steps=None, warmup=1):
    from oracle import lego_oracle as O
    torch.set_num_threads(os.cpu_count())
    shapes = O.state_shapes('nrms', HIDDEN, ADDITIVE, world.embed_dim, world.n_words, world.n_cats)
    state = O.default_state(shapes, {'embedding_vocab_table.glove.embedding.weight': torch.from_numpy(world.word_table)})
    spec = O.ModelSpec('nrms', HEADS, {world.title_col: world.word_vocab, 'category': 'category'})
    opt = torch.optim.Adam([v for v in state.values() if v.requires_grad], lr=1e-3)
    O.DROPOUT = DROPOUT
    B = next(iter(batches[0]['history']['input_ids'].values())).shape[0]

    def step(i):
        opt.zero_grad()
        loss = O.forward(state, spec, batches[i % len(batches)])
        loss.backward()
        opt.step()
        return loss.item()

    for i in range(warmup):
        step(i)
    t0 = time.time()
    n = 0
    while True:
        step(warmup + n)
        n += 1
        if steps is not None and n >= steps:
            break
        if steps is None and (time.time() - t0 >= budget_s and n >= 2):
            break
    dt = time.time() - t0
    O.DROPOUT = 0.0
    return dict(value=n * B / dt, steps=n, seconds=dt, batch=B)


def main():
    args = parse()
    if os.environ.get('LK_BENCH_WATCHDOG'):      # debugging aid: dump every thread's stack if the run exceeds N seconds
        import faulthandler
        faulthandler.dump_traceback_later(float(os.environ['LK_BENCH_WATCHDOG']), exit=True)
    rank = int(os.environ.get('RANK', 0))
    world_size = int(os.environ.get('WORLD_SIZE', 1))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    workload_name = 'NRMS train step, synthetic MIND-small shape (65,238 items, 400k x 300 GloVe-shaped table, title 30, history 50, 1+4 candidates, hidden 256)'
    base_cfg = dict(workload=workload_name, batch_per_gpu=args.batch, global_batch=args.batch * max(world_size, 1),
                    items_per_step_per_gpu=args.batch * (1 + NEG + WORKLOAD['hist_len']), dropout=DROPOUT,
                    parallelism=f'dp{world_size}' if world_size > 1 else 'single',
                    driver='autograd' if args.autograd else 'native (lk_nrms_fwd_bwd)')

    # ---------------- reference arm: CPU only, rank 0 only ----------------------------------------------------
    if args.impl == 'reference':
        if rank != 0:
            return
        from legommenders_b200.synth import MindWorld  # data generator only (numpy)
        world = build_world(args.small)
        batches = host_batches_cpu(world, args.batch, 4)
        bsz = args.batch
        r = cpu_reference(world, batches, steps=args.steps, warmup=max(1, min(args.warmup, 2)))
        base_cfg = dict(base_cfg, driver='reference CPU path restated on torch CPU (oracle/lego_oracle.py), all host threads', parallelism='cpu')
        line = dict(metric='NRMS train impressions/s', value=r['value'], unit='impressions/s', n_gpus=args.gpus, steps=args.steps,
                    warmup=args.warmup, ms_per_step=1000 * r['seconds'] / r['steps'], higher_is_better=True, scaling='weak',
                    vs_baseline=None, dtype='f32', data='synthetic', impl='reference', config=base_cfg,
                    cpu_baseline=dict(value=r['value'], unit='impressions/s', cores=os.cpu_count(), kind='port',
                                      sample=f'{r["steps"]} steps x {bsz} impressions, fwd+bwd+Adam, dropout {DROPOUT}, torch CPU fp32'),
                    e2e=dict(value=r['value'], unit='impressions/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0))
        print(json.dumps(line))
        return

    # ---------------- B200 arm ---------------------------------------------------------------------------------------
    import torch.distributed as dist
    if world_size > 1:
        # the contract is ONE line on stdout, and NCCL_DEBUG output goes to stdout by default: send it to a per-process file instead and
        # replay it on STDERR at the end (rank 0), so that whoever set NCCL_DEBUG (the driver's rank check) still sees the communicator lines
        nccl_log = None
        if os.environ.get('NCCL_DEBUG') and not os.environ.get('NCCL_DEBUG_FILE'):
            nccl_log = f'/tmp/lk_nccl_{os.getpid()}.log'
            os.environ['NCCL_DEBUG_FILE'] = nccl_log
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    import __graft_entry__ as ge
    if rank == 0:
        ge.build()
    if world_size > 1:
        dist.barrier()
    from legommenders_b200 import Env, _lib, builder
    from legommenders_b200.batching import BatchBuilder, tree_bytes, tree_to_device
    from legommenders_b200.trainer import FlatAdam, NativeNRMSStep

    torch.manual_seed(2023)
    world = build_world(args.small)
    model, resampler, cfg = builder.build_model(world, 'nrms', hidden=HIDDEN, heads=HEADS, additive=ADDITIVE, dropout=DROPOUT,
                                                neg_count=NEG, device_index=local_rank)
    dev = Env.device
    opt = FlatAdam(model, lr=1e-3)
    Env.train()
    model.train()

    bb = BatchBuilder(resampler, world, neg_count=NEG, seed=1000 + rank)
    rng = np.random.default_rng(77 + rank)
    POOL = 8
    host = [bb.train_batch(rng.integers(0, world.n_train, size=args.batch)) for _ in range(POOL)]
    devb = [tree_to_device(b, dev, non_blocking=False) for b in host]
    h2d = tree_bytes(host[0])

    native = None if args.autograd else NativeNRMSStep(model, opt)
    parity = None
    if world_size > 1 and native is not None and not args.no_parity_check:
        try:
            parity = parity_self_check(dev, rank, world_size, model, native, opt, resampler, world, args.batch)
        except Exception as e:   # noqa: BLE001 — reported in the line, never hidden
            parity = f'self-check raised {type(e).__name__}: {e}'[:300]
        flag = torch.tensor([0 if parity == 'ok' else 1], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if flag.item() and parity == 'ok':
            parity = 'another rank reported a mismatch'

    def step(batch):
        """One training step.  Default: the native driver (ONE C-ABI call enqueues forward+backward, then allreduce + Adam);
        --autograd: the same kernels driven by torch.autograd through the plugin surface (model(batch) / loss.backward())."""
        if native is not None:
            return native.step(batch)          # every gradient slot is overwritten by the driver: no zero_grad needed
        opt.zero_grad()
        loss = model(batch=batch)
        loss.backward()
        opt.step()
        return loss

    def sync_all():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k):
        """K steps bracketed by barrier + synchronize; device time by CUDA events; max over ranks.
        Also returns the per-step device times (an event after every step) so that a hiccup shows up as max >> median."""
        sync_all()
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(k + 1)]
        w0 = time.time()
        evs[0].record()
        for i in range(k):
            fn(i)
            evs[i + 1].record()
        sync_all()
        w1 = time.time()
        ms = evs[0].elapsed_time(evs[k])
        per = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(k))
        if world_size > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms, w0, w1, dict(median=per[len(per) // 2], min=per[0], max=per[-1])

    # The product's data path (batching.DeviceResampler): a step starts from B impression indices; negative sampling, history
    # concatenation and offsets are ONE kernel over device-resident tables (lk_resample_batch, enqueued one step ahead on a side stream so
    # that the four integers the host needs to size the step are back before it starts), token ids are expanded by lk_pack_item_tokens.
    #   value : indices of all timed steps resident in HBM before the clock starts; nothing is memoised — sampling and packing run every step
    #   e2e   : the same loop with the indices in pinned host memory (H2D of 8 B per impression inside the timed region, through the public
    #           DeviceResampler.submit / take + NativeNRMSStep.step calls) and a device->host read of the loss every step
    sampler = ClockSampler(local_rank) if rank == 0 else None   # started before the warm-up: its start-up is not timed
    for i in range(max(args.warmup, 3, POOL)):
        step(devb[i % POOL])
    n_loop = args.steps + 4
    if world_size > 1:
        # the GLOBAL batch of every step is drawn from one stream shared by all ranks and dealt out so that every rank packs (nearly) the
        # same number of token rows (batching.balanced_partition): a data-parallel step waits for its slowest rank
        from legommenders_b200.batching import balanced_partition, impression_costs
        g_rng = np.random.default_rng(4242)
        item_len_host = np.asarray([int(t['attention_mask'].sum()) for t in resampler.item_cache], dtype=np.int64)
        cost = impression_costs(world, item_len_host, NEG)
        rows_host = np.empty((n_loop, args.batch), dtype=np.int64)
        for i in range(n_loop):
            g_rows = g_rng.integers(0, world.n_train, size=world_size * args.batch)
            rows_host[i] = g_rows[balanced_partition(cost[g_rows], world_size)[rank]]
    else:
        rows_host = rng.integers(0, world.n_train, size=(n_loop, args.batch))
    if native is not None:
        from legommenders_b200.batching import DeviceResampler
        dres = DeviceResampler(resampler, world, dev, neg_count=NEG, seed=3000 + rank, max_batch=args.batch)
        rows_dev = torch.from_numpy(rows_host).to(dev)

        def resident_step(i):
            dres.submit(rows_dev[i + 1])
            return step(dres.take())

        loss_pin = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ev = [torch.cuda.Event() for _ in range(2)]
        losses_read = []

        def e2e_step(i):
            # every step's loss crosses to the host inside the timed region; the host READS it one step later (pinned buffer + event), so
            # that the read does not drain the device between steps — what a training loop that logs its loss does
            dres.submit(rows_host[i + 1])
            loss = step(dres.take())
            loss_pin[i & 1].copy_(loss, non_blocking=True)
            loss_ev[i & 1].record()
            if i > 0:
                loss_ev[(i - 1) & 1].synchronize()
                losses_read.append(float(loss_pin[(i - 1) & 1]))
            return loss

        for fn, src in ((resident_step, rows_dev), (e2e_step, rows_host)):       # warm both variants (allocator, pinned buffers)
            dres.submit(src[0])
            for i in range(3):
                fn(i)
            dres.take()
        dres.submit(rows_dev[0])
        l0 = _lib.load().lk_launch_count()
        if world_size > 1:
            opt.time_allreduce = []
        ms, w0, w1, per_step = timed(resident_step, args.steps)
        launches = int(_lib.load().lk_launch_count() - l0)
        collective = None
        if world_size > 1:
            ts = sorted(a.elapsed_time(b) for a, b in opt.time_allreduce)
            opt.time_allreduce = None
            collective = dict(op=('lk_allreduce_p2p (NVLink peer memory, ' + (('in-switch reduction via the multicast mapping, ' if opt._mc_ptr else '') + 'rendezvous inside the launch' if opt._flag_ptrs is not None else 'between two symmetric-memory barriers') + ')' if opt.symm is not None else 'ncclAllReduce') + ': sum of the flat gradient bucket', bytes=int(opt.grad.numel() * 4), ms_median=ts[len(ts) // 2], ms_min=ts[0],
                              ms_max=ts[-1], note='CUDA events on the step stream around the collective of every timed step, this rank: from local gradients '
                                                  'ready to reduced, i.e. the wait for the slowest rank + the transfer; it is not overlapped with compute')
        dres.take()
        dres.submit(rows_host[0])
        del losses_read[:]
        ms_e2e, _, _, per_step_e2e = timed(e2e_step, args.steps)       # timed() synchronises at the end: the last loss has landed too
        losses_read.append(float(loss_pin[(args.steps - 1) & 1]))
        assert len(losses_read) == args.steps and all(np.isfinite(losses_read)), losses_read
        dres.take()
        h2d_ids, d2h_e2e = args.batch * 8, 4 + 16        # indices up; loss + the resampler's four step-size integers down
    else:
        l0 = _lib.load().lk_launch_count()
        ms, w0, w1, per_step = timed(lambda i: step(devb[i % POOL]), args.steps)
        launches = int(_lib.load().lk_launch_count() - l0)
    clocks = sampler.stop(w0, w1) if sampler else None
    value = world_size * args.batch * args.steps / (ms / 1000.0)

    # the reference's wire format for comparison: the collated batch (int64 [B,55,S] trees, 3.7 MB) copied up and packed on the device
    def e2e_wire_step(i):
        b = tree_to_device(host[i % POOL], dev, non_blocking=True)
        return step(b).item()

    for i in range(2):
        e2e_wire_step(i)
    ms_wire, _, _, per_step_wire = timed(e2e_wire_step, args.steps)
    e2e_wire = dict(value=world_size * args.batch * args.steps / (ms_wire / 1000.0), unit='impressions/s', h2d_bytes_per_step=h2d,
                    d2h_bytes_per_step=4, ms_per_step=ms_wire / args.steps, ms_per_step_dist=per_step_wire)
    if native is None:
        ms_e2e, per_step_e2e, h2d_ids, d2h_e2e = ms_wire, per_step_wire, h2d, 4
    e2e_value = world_size * args.batch * args.steps / (ms_e2e / 1000.0)

    # strong scaling (SURVEY §8e asks for both): the GLOBAL batch fixed at 512 impressions, 512 / N per rank, same data path as `value`
    strong = None
    if native is not None and not args.no_extra and not args.small and 512 % world_size == 0:
        pb = 512 // world_size
        dres_s = DeviceResampler(resampler, world, dev, neg_count=NEG, seed=4000 + rank, max_batch=pb)
        rows_s = torch.from_numpy(rng.integers(0, world.n_train, size=(args.steps + 6, pb))).to(dev)

        def strong_step(i):
            dres_s.submit(rows_s[i + 1])
            return step(dres_s.take())

        dres_s.submit(rows_s[0])
        for i in range(3):
            strong_step(i)
        ms_s, _, _, per_s = timed(strong_step, args.steps)
        dres_s.take()
        strong = dict(scaling='strong', global_batch=512, batch_per_gpu=pb, n_gpus=world_size, ms_per_step=ms_s / args.steps,
                      impressions_per_s=512 * args.steps / (ms_s / 1000.0), ms_per_step_dist=per_s)
        del dres_s

    # per-entry-point device time over two extra steps (CUDA events around every C-ABI call on the launching stream)
    # (every rank runs the two steps — they contain the gradient all-reduce — but only rank 0 times them)
    roof, shares = None, None
    prof = None
    if rank == 0:
        if native is not None:
            _lib.native_profile_begin()
        else:
            prof = _lib.profile_begin()
    for i in range(2):
        step(devb[i % POOL])
    torch.cuda.synchronize()
    if rank == 0:
        shares, gemm = _lib.native_profile_end() if native is not None else _lib.profile_end(prof)
        pk = peaks()
        if gemm['ms'] > 0:
            ach = gemm['flops'] / (gemm['ms'] / 1e3) / 1e12
            tc = gemm['name'] == 'lk_tc_gemm'
            roof = dict(bound='tensor', kernel='tcgen05 contractions: tc_gemm_kernel + chain_kernel (tcgen05.mma, split-bf16 x3, fp32 TMEM accumulate)' if tc
                        else 'gemm_simt_kernel (fp32 FFMA)', achieved=ach, peak=pk['tensor'],
                        unit='TFLOP/s', frac=ach / pk['tensor'], traffic=measured_traffic('tcgen05_contractions') if tc else None,
                        traffic_note='dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over 40 consecutive tc_gemm_kernel / chain_kernel launches of a training step (profiles/r2_traffic.json, ncu --set full)',
                        pipe_tflops=3 * ach if tc else ach, pipe_frac=(3 * ach if tc else ach) / pk['tensor'], peak_source=pk['which'],
                        share_of_step=gemm['ms'] / max(sum(shares.values()), 1e-9), launches=gemm['calls'],
                        per_shape=gemm.get('per_shape'),
                        note='achieved = algorithmic flops (2*M*N*K per call, counted once) / CUDA-event time of the calls in a live step'
                             + ('; the kernel executes 3 bf16 MMAs per algorithmic product to hold fp32 parity, so the tensor pipe runs at 3x this rate' if tc else ''))

    line = dict(metric='NRMS train impressions/s', value=value, unit='impressions/s', n_gpus=world_size, steps=args.steps,
                warmup=args.warmup, ms_per_step=ms / args.steps, ms_per_step_dist=per_step, higher_is_better=True, scaling='weak', vs_baseline=None,
                dtype='f32', data='synthetic', impl='b200',
                config=dict(base_cfg, l2='per-step working set (~1 GB of activations) exceeds the 126 MB L2; every step samples a fresh batch (nothing memoised)',
                            inputs='impression indices resident in HBM; negative sampling, history concat and token packing run inside every timed step',
                            placement='global batch dealt to the ranks by packed-token cost (serpentine)' if world_size > 1 else 'single rank'),
                clocks=clocks,
                e2e=dict(value=e2e_value, unit='impressions/s', h2d_bytes_per_step=h2d_ids, d2h_bytes_per_step=d2h_e2e,
                         ms_per_step=ms_e2e / args.steps, ms_per_step_dist=per_step_e2e,
                         path='impression indices (pinned host) -> H2D -> lk_resample_batch (side stream, one step ahead) -> lk_pack_item_tokens -> lk_nrms_fwd_bwd -> allreduce -> Adam -> D2H loss (async copy every step, read by the host one step later)'
                         if native is not None else 'wire-format batch'),
                e2e_wire_format=e2e_wire,
                gpu_launches=launches, roofline=roof, kernel_ms_share=shares, strong_scaling=strong)
    if native is not None and world_size > 1:
        line['collective'] = collective

    if parity is not None:
        line['parity_check'] = parity
    if world_size > 1 and not args.small and not args.no_extra:
        extra = {}
        for name, fn in (('config4_sharded_table_train', lambda: config4_line(dev, rank, world_size, args.batch)),
                         ('cached_eval_sharded', lambda: sharded_eval_line(dev, rank, world_size, model, resampler, world)),
                         ('llm_item_path_sharded', lambda: config5_line(dev, rank, world_size, peaks()['hbm'], peaks()['tensor']))):
            try:                                                             # collective: every rank takes part
                extra[name] = fn()
            except Exception as e:   # noqa: BLE001
                extra[name] = dict(error=f'{type(e).__name__}: {e}'[:300])
        if rank == 0:
            line['extra'] = extra
    if rank == 0 and world_size == 1 and not args.small and not args.no_extra:
        line['extra'] = hbm_bound_lines(dev, world, peaks()['hbm'], model=model, resampler=resampler, tensor_peak=peaks()['tensor'])
    if rank == 0 and world_size == 1 and not args.no_cpu_baseline:
        hb = [{k: v for k, v in b.items()} for b in host[:4]]
        r = cpu_reference(world, hb, budget_s=args.cpu_seconds)
        line['cpu_baseline'] = dict(value=r['value'], unit='impressions/s', cores=os.cpu_count(), kind='port',
                                    sample=f'{r["steps"]} steps x {r["batch"]} impressions ({r["seconds"]:.1f} s), fwd+bwd+Adam, dropout {DROPOUT}, torch CPU fp32')
    if rank == 0:
        print(json.dumps(line))
    if world_size > 1:
        dist.barrier()
        dist.destroy_process_group()
        if rank == 0 and nccl_log and os.path.exists(nccl_log):
            with open(nccl_log) as f:
                sys.stderr.write(f.read())


def parity_self_check(dev, rank, W, model, native, opt, resampler, world, batch):
    """N > 1 only: the multi-rank correctness checks of tests/test_gpu_multirank.py, run on the box that is about to be timed (the driver's
    GPU-test box has one GPU).  (1) data-parallel gradients (local batch, all-reduce mean) == gradients of the GLOBAL batch computed by one
    rank; (2) row-sharded table lookup bit-identical to indexing the replicated table; (3) sharded cached evaluation (user caches and rows
    partitioned by user id, item cache all-gathered) gives the metrics of the unsharded evaluation.  -> 'ok' or a description."""
    import torch.distributed as dist
    from legommenders_b200 import Env, evaluate as ev, sharding
    from legommenders_b200.batching import DeviceBatcher, DeviceResampler
    problems = []
    # (1) ---------------------------------------------------------------------------------------------------------------
    B = min(batch, 32)
    dres = DeviceResampler(resampler, world, dev, neg_count=NEG, seed=9, max_batch=B * W)
    all_rows = np.arange(B * W) % world.n_train
    dres.submit(all_rows[rank * B:(rank + 1) * B])
    native.fwd_bwd(dres.take(), training=False)
    g_dp = opt.grad.clone()
    dist.all_reduce(g_dp, op=dist.ReduceOp.SUM)
    if opt.symm is not None:                                      # the timed step's own collective (lk_allreduce_p2p) against ncclAllReduce
        opt.allreduce()
        err = float((opt.grad - g_dp).abs().max() / g_dp.abs().max())
        if not err <= 1e-6:
            problems.append(f'peer-memory all-reduce differs from ncclAllReduce: {err:.2e}')
        r0 = opt.grad.clone()
        dist.broadcast(r0, src=0)
        if not torch.equal(r0, opt.grad):
            problems.append('peer-memory all-reduce: ranks hold different sums')
    g_dp /= W
    dres.submit(all_rows)
    loss_g = native.fwd_bwd(dres.take(), training=False).item()
    g_glob = opt.grad.clone()
    err = float((g_dp - g_glob).abs().max() / g_glob.abs().max())
    if not err <= 5e-5:
        problems.append(f'dp gradients differ from the global-batch gradients: {err:.2e}')
    # (2) ---------------------------------------------------------------------------------------------------------------
    V, E = 200_000, 300
    gen = torch.Generator(device=dev).manual_seed(123)           # same table on every rank
    full = torch.empty((V, E), dtype=torch.float32, device=dev).normal_(0, 0.4, generator=gen)
    st = sharding.ShardedTable(sharding.shard_rows(full, rank, W), V)
    ids = torch.randint(-1, V, (50_000,), generator=torch.Generator(device=dev).manual_seed(1000 + rank), device=dev)
    rows, inv = st.lookup_unique(ids)
    got = rows[inv.clamp(min=0)] * (inv >= 0).unsqueeze(1)
    want = full[ids.clamp(min=0)] * (ids >= 0).unsqueeze(1)
    if not torch.equal(got, want):
        problems.append('sharded lookup is not bit-identical to the replicated table')
    del full, st
    # (3) ---------------------------------------------------------------------------------------------------------------
    Env.test(); model.eval()
    dbat = DeviceBatcher(resampler, world, dev)
    eu, ei, el = (torch.from_numpy(a) for a in (world.eval_users, world.eval_items, world.eval_click))
    ev.build_caches_device(model, dbat)                           # sharded: item slices all-gathered, users by id % W
    sharded, _, _ = ev.evaluate(model, eu, ei, el)
    item_repr = model.cacher.item.repr
    user_parts = model.cacher.user.repr.clone()
    dist.all_reduce(user_parts, op=dist.ReduceOp.SUM)             # every user row was filled by exactly one rank (zeros elsewhere)
    model.cacher.user.repr = user_parts
    full_vals, _, _ = ev.evaluate(model, eu, ei, el, shard=False)
    for k in sharded:
        if abs(sharded[k] - full_vals[k]) > 2e-6:
            problems.append(f'sharded evaluation {k}: {sharded[k]} vs {full_vals[k]}')
    model.cacher.clean()
    Env.train(); model.train()
    return 'ok' if not problems else '; '.join(problems)


def sharded_eval_line(dev, rank, W, model, resampler, world):
    """Config 3 (BASELINE.json metric: cached-eval scores/s): representation caches built from id lists on the device (items sliced over
    the ranks and all-gathered, users partitioned by user id), then every validation row scored by the rank that owns its user and the
    five default metrics reduced with one all-reduce.  Device-timed, max over ranks; ids start in pinned host memory."""
    import torch.distributed as dist
    from legommenders_b200 import Env, evaluate as ev
    from legommenders_b200.batching import DeviceBatcher
    Env.test(); model.eval()
    dbat = DeviceBatcher(resampler, world, dev)
    ev.build_caches_device(model, dbat)                           # warm-up
    torch.cuda.synchronize()
    if W > 1:
        dist.barrier()
    t0 = time.time()
    ev.build_caches_device(model, dbat)
    torch.cuda.synchronize()
    if W > 1:
        dist.barrier()
    build_s = time.time() - t0
    from legommenders_b200 import sharding
    R = int(len(world.eval_users))
    own = sharding.owned_rows(torch.from_numpy(world.eval_users), rank, W)      # the evaluation set is partitioned by user id ONCE (layout, untimed)
    hu, hi_, hl = (torch.from_numpy(a)[own].pin_memory() for a in (world.eval_users, world.eval_items, world.eval_click))
    res = {}
    for _ in range(2):
        res['m'] = ev.evaluate(model, hu, hi_, hl, presharded=True)[0]
    torch.cuda.synchronize()
    if W > 1:
        dist.barrier()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        res['m'] = ev.evaluate(model, hu, hi_, hl, presharded=True)[0]
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    if W > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    model.cacher.clean()
    Env.train(); model.train()
    return dict(rows=R, rows_this_rank=int(hu.numel()), n_gpus=W, ms=ms, scores_per_s=R / ms * 1e3, cache_build_seconds=build_s,
                items=int(world.n_items), users=int(world.n_users), metrics={k: round(float(v), 4) for k, v in res['m'].items()},
                note='evaluate(presharded): pinned host ids of the rows whose user this rank owns -> H2D -> lk_cached_scores -> lk_group_metrics -> all-reduce of (sum, count); '
                     'cache build = evaluate.build_caches_device (wall clock, incl. the item-cache all-gather)')


def config4_line(dev, rank, W, batch, steps=20):
    """Config 4 (BASELINE.json): NRMS training with a 4M x 300 fp32 word table ROW-SHARDED over the ranks (row i on rank i % W) and
    history length 100.  Every step: lk_resample_batch -> lk_pack_item_tokens -> sharded lookup of the batch's distinct title tokens
    (lk_shard_plan: dedup + owner bucketing on the device, fixed-capacity NCCL all-to-all of ids, lk_shard_gather on the owner, all-to-all
    of rows) -> lk_nrms_fwd_bwd on the compact table -> gradient all-reduce -> Adam.  Device-timed, max over ranks; the lookup is also
    timed alone on the same batches."""
    import torch.distributed as dist
    from legommenders_b200 import Env, builder, sharding
    from legommenders_b200.batching import DeviceResampler
    from legommenders_b200.synth import MindWorld
    from legommenders_b200.trainer import FlatAdam, NativeNRMSStep
    V, E, H = 4_000_000, 300, 100
    w4 = MindWorld(seed=4242, make_table=False, **dict(WORKLOAD, n_words=V, hist_len=H, n_train=50_000, n_eval_groups=8, eval_group_mean=4))
    # the model's own (replicated) table parameter is an UNINITIALISED device placeholder of the right shape: with a sharded table the step
    # never reads it (trainer.NativeNRMSStep), it only satisfies the EmbeddingHub's vocabulary-size check
    w4.word_table = torch.empty((V, E), dtype=torch.float32, device=dev)
    torch.manual_seed(77)
    model4, res4, _ = builder.build_model(w4, 'nrms', hidden=HIDDEN, heads=HEADS, additive=ADDITIVE, dropout=DROPOUT, neg_count=NEG,
                                          device_index=dev.index)
    local = torch.empty(((V - rank + W - 1) // W, E), dtype=torch.float32, device=dev).normal_(0, 0.4)
    st = sharding.ShardedTable(local, V)
    opt4 = FlatAdam(model4, lr=1e-3)
    nat4 = NativeNRMSStep(model4, opt4, sharded_table=st)
    Env.train(); model4.train()
    dres = DeviceResampler(res4, w4, dev, neg_count=NEG, seed=500 + rank, max_batch=batch)
    rng = np.random.default_rng(900 + rank)
    rows = torch.from_numpy(rng.integers(0, w4.n_train, size=(steps + 8, batch))).to(dev)

    def one(i):
        dres.submit(rows[i + 1])
        return nat4.step(dres.take())

    def sync():
        torch.cuda.synchronize()
        if W > 1:
            dist.barrier()
            torch.cuda.synchronize()

    dres.submit(rows[0])
    for i in range(5):
        one(i)
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(5, 5 + steps):
        loss = one(i)
    e1.record()
    sync()
    st.check()
    last = dres.take()
    t = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    # the lookup alone, on the last batch's title tokens
    pk = last['__lk_packed__'][0]
    title = pk.ids[nat4.title_col]
    for _ in range(3):
        st.lookup_unique(title)
    sync()
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    for _ in range(steps):
        tab, inv = st.lookup_unique(title)
    l1.record()
    sync()
    t2 = torch.tensor([l0.elapsed_time(l1) / steps], device=dev)
    uniq = torch.tensor([float(torch.unique(title[title >= 0]).numel())], device=dev)
    if W > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        dist.all_reduce(uniq, op=dist.ReduceOp.SUM)
    ms, ms_l, u = t.item(), t2.item(), uniq.item() / W
    out = dict(workload='NRMS train, 4M x 300 word table row-sharded over the ranks, history 100, 1+4 candidates, hidden 256', n_gpus=W,
               batch_per_gpu=batch, impressions_per_s=W * batch / ms * 1e3, ms_per_step=ms, loss=float(loss.item()),
               token_rows_last_batch=int(pk.rows), items_last_batch=int(pk.n), ms_per_lookup=ms_l, unique_rows_per_rank=u, bucket_capacity=st._cap,
               nvlink_GBps_per_rank=(W - 1) * st._cap * E * 4 / ms_l / 1e6 if W > 1 else 0.0,
               useful_nvlink_GBps_per_rank=u * (W - 1) / W * E * 4 / ms_l / 1e6 if W > 1 else 0.0,
               note='lookup = lk_shard_plan + lk_shard_inverse + all-to-all(ids, fixed capacity) + lk_shard_gather + all-to-all(rows): no sort, no split '
                    'sizes, no host round trip; rows cross NVLink once per distinct token; user encoder histories up to 100 run the SIMT attention kernels')
    del model4, opt4, nat4, st, local, w4.word_table
    torch.cuda.empty_cache()
    return out


def config5_line(dev, rank, W, hbm_peak, tensor_peak, n_items=1_000_000, E_llm=4096, U_sw=4096, k=10):
    """Config 5 (BASELINE.json): frozen [1M, 4096] LLM item embeddings ROW-SHARDED over the ranks (contiguous blocks), projected locally to
    256-d on the tensor cores, then the catalog sweep of 4096 replicated users with the per-user top-k fused into the epilogue
    (lk_sweep_topk: no U x N score matrix exists) and one all-gather + k-way merge of the [U, k] candidates.  Device-timed, max over ranks."""
    import torch.distributed as dist
    from legommenders_b200 import ops, sharding
    D = HIDDEN
    a, b = sharding.item_slice(n_items, rank, W)
    n_loc = b - a
    g = torch.Generator(device=dev).manual_seed(55 + rank)
    tab = torch.empty((n_loc, E_llm), dtype=torch.float32, device=dev).normal_(0, 1, generator=g)
    gw = torch.Generator(device=dev).manual_seed(7)                  # replicated projection and users: same on every rank
    Wp_f = torch.empty((D, E_llm), dtype=torch.float32, device=dev).normal_(0, 0.02, generator=gw)
    bias = torch.zeros(D, dtype=torch.float32, device=dev)
    users = torch.empty((U_sw, D), dtype=torch.float32, device=dev).normal_(0, 1, generator=gw)

    def timed(fn, reps=3):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        if W > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
        if W > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    res = {}
    ms_split = timed(lambda: res.__setitem__('tp', ops.split_planes(tab)), reps=2)      # fp32 .npy rows (embed.py) -> operand planes
    tp = res.pop('tp')
    ref_rows = torch.arange(0, n_loc, max(1, n_loc // 64), device=dev)[:64]
    ref = (tab[ref_rows].double() @ Wp_f.double().t()).float()
    del tab
    Wp = ops.split_planes(Wp_f)
    rep = torch.empty((n_loc, D), dtype=torch.float32, device=dev)
    ms_p = timed(lambda: ops.tc_gemm(tp, Wp, False, n_loc, D, E_llm, out=rep, bias=bias))
    err = float((rep[ref_rows] - ref).abs().max() / ref.abs().max())
    del tp
    Ip, Up = ops.split_planes(rep), ops.split_planes(users)
    ms_s = timed(lambda: res.__setitem__('tk', sharding.catalog_topk(Up, Ip, k, item_offset=a)))
    vals, idx = res['tk']
    # spot check of the fused top-k against explicit scores for a few users (local shard only when W > 1 would miss winners: use the merged result
    # against an all-gathered brute force on 8 users)
    probe = users[:8]
    loc = probe @ rep.t()
    lv, li = torch.topk(loc, k, dim=1)
    li = li + a
    if W > 1:
        gv = [torch.empty_like(lv) for _ in range(W)]; gi = [torch.empty_like(li) for _ in range(W)]
        dist.all_gather(gv, lv.contiguous()); dist.all_gather(gi, li.contiguous())
        lv, li = ops.merge_topk(torch.cat(gv, 1), torch.cat(gi, 1), k)
    agree = float((idx[:8] == li).double().mean())
    fl_p, fl_s = 2.0 * n_items * D * E_llm, 2.0 * U_sw * n_items * D
    by_p = n_items * E_llm * 4 + n_items * D * 4
    by_s = n_items * D * 4 + U_sw * D * 4                       # the sweep reads every projected item row once, algorithmically
    return dict(items=n_items, items_per_rank=n_loc, embed_dim=E_llm, users=U_sw, k=k, n_gpus=W,
                split_ms=ms_split, projection_ms=ms_p, projection_incl_split_ms=ms_split + ms_p,
                projection=dict(items_per_s=n_items / ms_p * 1e3, items_per_s_incl_split=n_items / (ms_p + ms_split) * 1e3,
                                tflops=fl_p / ms_p / 1e9 / 1.0, tensor_frac=fl_p / W / ms_p / 1e9 / tensor_peak,
                                pipe_tensor_frac=3 * fl_p / W / ms_p / 1e9 / tensor_peak, hbm_frac=by_p / W / ms_p / 1e6 / hbm_peak,
                                max_rel_err_vs_fp64=err),
                sweep_topk=dict(ms=ms_s, scores_per_s=U_sw * n_items / ms_s * 1e3, tflops=fl_s / ms_s / 1e9,
                                tensor_frac=fl_s / W / ms_s / 1e9 / tensor_peak, pipe_tensor_frac=3 * fl_s / W / ms_s / 1e9 / tensor_peak,
                                top_k_agreement_with_brute_force=agree,
                                note='per-rank tensor fractions; scores are never written: the 16.8 GB fp32 score matrix of the unfused sweep does not exist'),
                note='lk_tc_gemm projection (tcgen05 split-bf16 x3) + lk_sweep_topk; tflops are whole-job algorithmic (the pipe executes 3 MMAs per product)')


def hbm_bound_lines(dev, world, hbm_peak, model=None, resampler=None, tensor_peak=1400.0):
    """The HBM-bound pieces of the metric (BASELINE.json: cached-eval scores/s, gather HBM GB/s), each timed alone with CUDA events
    (5 repetitions after 2 warm-ups, inputs resident, a 256 MB write between repetitions to flush L2) and set against the measured
    copy bandwidth.  Algorithmic bytes per unit as in DESIGN.md §4."""
    from legommenders_b200 import ops
    from legommenders_b200.metrics import MetricPool
    g = torch.Generator(device='cpu').manual_seed(5)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def time_it(fn, reps=5):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return sorted(ts)[len(ts) // 2]

    out = {}

    def guarded(name):
        """Each extra line is independent: a failure is recorded under its own key and the others still run."""
        def deco(fn):
            try:
                fn()
            except Exception as e:   # noqa: BLE001
                out[name] = dict(error=f'{type(e).__name__}: {e}'[:300])
            try:
                torch.cuda.synchronize()
            except Exception as e:   # noqa: BLE001 — the bench line must still be printed
                out.setdefault(name, {})['sync_error'] = f'{type(e).__name__}: {e}'[:200]
            return fn
        return deco

    # (5) cached evaluation at MIND-small validation shape: every impression row against the two caches
    D = HIDDEN
    U = torch.randn(world.n_users, D, generator=g).to(dev)
    I = torch.randn(world.n_items, D, generator=g).to(dev)
    R = int(len(world.eval_users))
    uid, iid = torch.from_numpy(world.eval_users).to(dev), torch.from_numpy(world.eval_items).to(dev)
    lab = torch.from_numpy(world.eval_click).to(dev)
    sc = torch.empty(R, dtype=torch.float32, device=dev)
    ms = time_it(lambda: ops.cached_scores(U, I, uid, iid, out=sc))
    by = R * (2 * D * 4 + 2 * 8 + 4)
    out['cached_eval'] = dict(rows=R, ms=ms, scores_per_s=R / ms * 1e3, algorithmic_GBps=by / ms / 1e6, bytes_per_score=2 * D * 4 + 20,
                              l2_resident=True,
                              note='rows sorted by user (as the evaluation set is stored): the user row is reused ~36x and the 67 MB item cache fits the L2 '
                                   '(ncu: 76 % L2 hit, 0.42 GB of DRAM reads for 5.4 GB algorithmic, profiles/r2_03) - NOT an HBM roofline figure')
    perm = torch.randperm(R, generator=g).to(dev)
    uid_s = uid[perm].contiguous()
    ms_sh = time_it(lambda: ops.cached_scores(U, I, uid_s, iid, out=sc))
    out['cached_eval_shuffled'] = dict(rows=R, ms=ms_sh, scores_per_s=R / ms_sh * 1e3, algorithmic_GBps=by / ms_sh / 1e6,
                                       hbm_frac=by / ms_sh / 1e6 / hbm_peak,
                                       note='user ids shuffled: both caches (161 MB) exceed the L2 - the HBM-bound variant of the same kernel; algorithmic bytes '
                                            '2,068 per score (ncu: 2.6 GB of DRAM reads, 39 % L2 hit)')
    ops.cached_scores(U, I, uid, iid, out=sc)
    pool = MetricPool.parse(['GAUC', 'MRR', 'NDCG@1', 'NDCG@5', 'NDCG@10'])
    ms_m = time_it(lambda: pool.calculate(sc, lab, uid), reps=3)
    out['group_metrics'] = dict(rows=R, groups=pool.n_groups, ms=ms_m, rows_per_s=R / ms_m * 1e3,
                                note='GAUC, MRR, NDCG@1/5/10 incl. the sort by group key and the D2H of the 5 means')
    # (1) gather + masked mean pooling: MIND-small items over the 400k x 300 table, then a 4M-row table (config 4 size, HBM-resident)
    E = world.embed_dim
    ids = torch.from_numpy(world.item_title_matrix()).to(dev)                     # [N_items, 30], -1 padded
    table = torch.from_numpy(world.word_table).to(dev)
    N, S = ids.shape
    real = int((ids > -1).sum().item())
    ms = time_it(lambda: ops.gather_pool(ids, None, table))
    by = real * E * 4 + N * S * 8 + N * E * 4
    out['gather_pool_mind_small'] = dict(items=N, real_tokens=real, ms=ms, algorithmic_GBps=by / ms / 1e6, hbm_frac=by / ms / 1e6 / hbm_peak,
                                         note='Zipf token ids: the hot rows live in L2')
    big = torch.empty((4_000_000, E), dtype=torch.float32, device=dev).normal_(0, 0.4)
    ids_b = torch.randint(0, big.shape[0], (N * 4, S), generator=g).to(dev)
    ids_b[:, 24:] = -1
    real = int((ids_b > -1).sum().item())
    ms = time_it(lambda: ops.gather_pool(ids_b, None, big))
    by = real * E * 4 + ids_b.numel() * 8 + ids_b.shape[0] * E * 4
    out['gather_pool_4M_rows'] = dict(items=ids_b.shape[0], real_tokens=real, ms=ms, algorithmic_GBps=by / ms / 1e6,
                                      hbm_frac=by / ms / 1e6 / hbm_peak, note='4.8 GB table, uniform ids: HBM-resident rows of 1200 B')
    del big, ids_b, table, ids

    # (5, end to end) config 3 through the public evaluate() call: pinned host ids -> H2D -> lk_cached_scores -> lk_group_metrics ->
    # D2H of the metric means, every repetition.  The two caches are attached to the model's cacher the way ReprCacher leaves them.
    @guarded('cached_eval_e2e')
    def _():
        from legommenders_b200 import evaluate as ev
        cacher = model.cacher
        cacher.item.repr, cacher.user.repr = I, U
        cacher.item.cached = cacher.user.cached = True
        hu, hi_, hl = (torch.from_numpy(a).pin_memory() for a in (world.eval_users, world.eval_items, world.eval_click))
        res = {}

        def run():
            res['m'] = ev.evaluate(model, hu, hi_, hl)[0]
        ms_e = time_it(run, reps=3)
        cacher.item.repr = cacher.user.repr = None
        cacher.item.cached = cacher.user.cached = False
        out['cached_eval_e2e'] = dict(rows=R, ms=ms_e, scores_per_s=R / ms_e * 1e3, h2d_bytes=int(3 * 8 * R), d2h_bytes=5 * 8,
                                      metrics={k: round(float(v), 4) for k, v in res['m'].items()},
                                      note='evaluate(): H2D of user/item/label ids (int64, as the reference holds them; the group key is the user '
                                           'id and is copied once), scoring, GAUC/MRR/NDCG@1,5,10, D2H of the means')
        cacher.item.repr, cacher.user.repr = I, U
        cacher.item.cached = cacher.user.cached = True
        hu32, hi32, hl32 = (torch.from_numpy(a.astype(np.int32)).pin_memory() for a in (world.eval_users, world.eval_items, world.eval_click))

        def run32():
            res['m32'] = ev.evaluate(model, hu32, hi32, hl32)[0]
        ms_32 = time_it(run32, reps=3)
        cacher.item.repr = cacher.user.repr = None
        cacher.item.cached = cacher.user.cached = False
        out['cached_eval_e2e_int32'] = dict(rows=R, ms=ms_32, scores_per_s=R / ms_32 * 1e3, h2d_bytes=int(3 * 4 * R), d2h_bytes=5 * 8,
                                            metrics_equal=bool(res['m32'] == res['m']),
                                            note='the same call with the evaluation ids held as int32 in pinned host memory: half the PCIe '
                                                 'bytes, widened on the device')

    # (a15) item-representation cache build through the cacher contract (positional content list, pages of cache_page_size)
    @guarded('item_cache_build')
    def _():
        from legommenders_b200 import Env
        Env.test(); model.eval()
        contents = resampler.item_cache
        model.cacher.item.cache(contents[:2048])
        torch.cuda.synchronize()
        t0 = time.time()
        model.cacher.item.cache(contents)
        torch.cuda.synchronize()
        dt = time.time() - t0
        model.cacher.clean()
        Env.train(); model.train()
        out['item_cache_build'] = dict(items=len(contents), seconds=dt, items_per_s=len(contents) / dt,
                                       note='ItemCacher.cache(resampler.item_cache): host stacking of per-item id rows (Python) + packed NRMS item encoder per page, wall clock')

    # (a15, device side) both caches from id lists only: evaluate.build_caches_device
    @guarded('cache_build_device')
    def _():
        from legommenders_b200 import Env, evaluate as ev
        from legommenders_b200.batching import DeviceBatcher
        Env.test(); model.eval()
        dbat = DeviceBatcher(resampler, world, dev)
        ev.build_caches_device(model, dbat)                  # warm-up (allocator, kernel attributes)
        torch.cuda.synchronize()
        t0 = time.time()
        ev.build_caches_device(model, dbat)
        torch.cuda.synchronize()
        dt = time.time() - t0
        n_i, n_u = int(model.cacher.item.repr.shape[0]), len(dbat.hist)
        model.cacher.clean()
        Env.train(); model.train()
        out['cache_build_device'] = dict(items=n_i, users=n_u, seconds=dt, items_and_users_per_s=(n_i + n_u) / dt,
                                         note='item pages: lk_pack_item_tokens + packed NRMS item encoder; user pages: lk_index_rows over the item cache + packed NRMS user encoder; wall clock incl. host offset arithmetic')

    # config 1 (BASELINE.json configs[0]: NAML, the reference's CPU-runnable case) and the neighbouring model families (LSTUR, MINER,
    # Fastformer) on the GPU through the plugin surface + torch.autograd (model(batch); loss.backward(); FlatAdam.step()), pre-built device batches
    def plugin_train_line(kind, note):
        from legommenders_b200 import Env, builder
        from legommenders_b200.batching import BatchBuilder, tree_to_device
        from legommenders_b200.trainer import FlatAdam
        torch.manual_seed(11)
        m2, r2, _ = builder.build_model(world, kind, hidden=HIDDEN, heads=HEADS, additive=ADDITIVE, dropout=DROPOUT, neg_count=NEG,
                                        device_index=dev.index)
        o2 = FlatAdam(m2, lr=1e-3)
        Env.train(); m2.train()
        bb2 = BatchBuilder(r2, world, neg_count=NEG, seed=3)
        rng2 = np.random.default_rng(8)
        pool2 = [tree_to_device(bb2.train_batch(rng2.integers(0, world.n_train, size=64)), dev, non_blocking=False) for _ in range(4)]

        def one(i):
            o2.zero_grad()
            loss = m2(batch=pool2[i % 4])
            loss.backward()
            o2.step()
            return loss

        for i in range(4):
            one(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K2 = 10
        e0.record()
        for i in range(K2):
            loss = one(i)
        e1.record()
        torch.cuda.synchronize()
        ms2 = e0.elapsed_time(e1) / K2
        del m2, o2, pool2
        return dict(ms_per_step=ms2, impressions_per_s=64 / ms2 * 1e3, loss=float(loss.item()), batch=64, note=note)

    for kind, note in (('naml', 'NAML: CNN title encoder (conv forward on fp32 FFMA lk_conv1d_fwd, input and weight gradients as tcgen05 contractions over im2col planes) + additive attention, Ada user encoder, DotPredictor + CE'),
                       ('lstur', 'LSTUR: CNNCat item encoder + GRU user encoder (lk_gru_fwd/bwd)'),
                       ('miner', 'MINER: 2-layer BERT-style Transformer item encoder + poly attention (8 codes) + target-aware predictor'),
                       ('fastformer', 'Fastformer item and user encoders (1 layer each)')):
        @guarded(kind + '_train')
        def _(kind=kind, note=note):
            out[kind + '_train'] = plugin_train_line(kind, note + '; generic plugin path, host-driven launches, MIND-small shape')

    @guarded('naml_train_tc_forward')
    def _():
        from legommenders_b200 import ops as _ops
        old = _ops.CONV_TC_FWD
        _ops.CONV_TC_FWD = True
        try:
            out['naml_train_tc_forward'] = plugin_train_line('naml', 'NAML with LK_CONV_TC_FWD=1: the conv forward on the tensor cores as well (not the parity '
                                                             'configuration: ReLU gates within 1e-5 of zero may flip, see ops._Conv1dReluMask)')
        finally:
            _ops.CONV_TC_FWD = old

    # config 4 on one GPU (the row-sharded table with a single shard: same kernels, no NVLink traffic); the N > 1 lines are in SCALE
    @guarded('config4_sharded_table_train')
    def _():
        out['config4_sharded_table_train'] = config4_line(dev, 0, 1, 64)

    # (2) LLM-embedding item path (config 5) on one GPU; the row-sharded N > 1 lines are in SCALE
    @guarded('llm_item_path')
    def _():
        out['llm_item_path'] = config5_line(dev, 0, 1, hbm_peak, tensor_peak)
    return out


def host_batches_cpu(world, batch, n):
    """Wire-format batches without touching CUDA (reference arm): inputer layouts via the oracle's integer restatement."""
    from collections import OrderedDict
    from oracle import lego_oracle as O
    inputs = [world.title_col, 'category']
    max_lens = {world.title_col: world.title_len, 'category': None}
    tab = world.item_table()
    cols = None
    rows = []
    for i in range(len(tab)):
        rows.append(O.concat_layout(tab[i], inputs, max_lens, False, True))
    ids = OrderedDict((c, torch.from_numpy(np.stack([r['input_ids'][c] for r in rows]))) for c in rows[0]['input_ids'])
    am = torch.from_numpy(np.stack([r['attention_mask'] for r in rows]))
    rng = np.random.default_rng(5)
    out = []
    for _ in range(n):
        r = rng.integers(0, world.n_train, size=batch)
        users = world.train_users[r]
        cand = np.concatenate([world.train_pos[r][:, None], rng.integers(0, world.n_items, size=(batch, NEG))], axis=1)
        hist = np.zeros((batch, world.hist_len), dtype=np.int64)
        hm = np.zeros((batch, world.hist_len), dtype=np.int64)
        for b, u in enumerate(users):
            h = world.histories[int(u)]
            hist[b, :len(h)] = h
            hm[b, :len(h)] = 1
        ct, ht = torch.from_numpy(cand), torch.from_numpy(hist)
        out.append({'item_id': dict(input_ids=OrderedDict((c, v[ct]) for c, v in ids.items()), attention_mask=am[ct]),
                    'history': dict(input_ids=OrderedDict((c, v[ht]) for c, v in ids.items()), attention_mask=am[ht]),
                    '__clicks_mask__': torch.from_numpy(hm), 'user_id': torch.from_numpy(users),
                    'click': torch.ones(batch, dtype=torch.int64)})
    return out


if __name__ == '__main__':
    main()

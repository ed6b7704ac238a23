#!/usr/bin/env python
"""Benchmark of the hot path: batched eval_jacob, states/s (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One *step* = one pass of eval_jacob over one batch of synthetic states of the GRI-Mech-3.0-
shaped mechanism (53 species / 325 reactions; the real GRI file is not available offline, see
DESIGN.md).  With N > 1 (launched by torchrun, one rank per GPU) every rank evaluates its own
batch of the same size -- the state batch is partitioned, there is no data-path collective --
and `value` is the states all ranks processed divided by the max-over-ranks device time.

value        device-resident: states and Jacobians stay in HBM in the reference's GPU layout
             (state-fastest / struct-of-arrays: y[NSP][n], jac[NSP*NSP][n], mech_auxiliary.py:418-420),
             CUDA events around K launches
e2e          the same metric through the host-pointer C-ABI call (pyjac_eval_jacob_host) with
             pinned HOST buffers: H2D of the states and D2H of every Jacobian inside the timing
roofline     HBM bound; algorithmic bytes = 8*NSP^2 + 8*(NSP+1) per state
cpu_baseline the reference's own generated C (oracle/_ref, OpenMP over states, all host
             threads) on a bounded sample of the same states; rank 0, N = 1 only

workloads    sub-records for BASELINE.json's other configurations on this rank count: USC-II-shaped
             (configs[2]), n-heptane-shaped at 32 768 states per GPU (configs[3]), the 1020 H2/O2 PaSR
             states (configs[0]); --no-workloads leaves them out
with_gather  (N > 1) the same batch with every Jacobian gathered to rank 0 over NCCL, chunked so that
             the send of chunk c overlaps the kernel of chunk c + 1; rank 0 receives into a bounded ring
             of buffers (8 x 23.6 GB would not fit one GPU).  `ingress_gbs` = bytes into rank 0 / time
strong       (N > 1) the fixed 2^20-state batch split over the N ranks

--impl reference times the reference alone (it has no working sm_100 GPU path): its generated C under its own
harness (tester.c.in + read_initial_conditions.c + timer.h, oracle/_ref/<name>/speedtest) on all host threads.
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

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MECH_FILE = os.path.join(ROOT, 'tests', 'golden', 'gri30_syn.inp')
REF_NAME = 'gri30'
WORKLOAD = 'GRI-3.0-shaped synthetic mechanism (53 sp / 325 rxn), eval_jacob, fp64'
# --workload: BASELINE.json configs[1] (default, the one the metric is quoted on), [2], [3];
# (mechanism shape of pyjac_b200/synth.py, name under oracle/_ref, description, states per GPU)
WORKLOADS = {
    'gri30': (None, 'gri30', WORKLOAD, 1 << 20),
    'usc2': ('usc2', 'usc2', 'USC-Mech-II-shaped synthetic mechanism (111 sp / 784 rxn), eval_jacob, fp64', 1 << 17),
    'nc7': ('nc7', 'nc7', 'n-heptane-shaped synthetic mechanism (654 sp / 2827 rxn), eval_jacob, fp64', 18944),
}


def select_workload(name: str):
    """Point MECH_FILE / REF_NAME / WORKLOAD at one of WORKLOADS; returns its default batch."""
    global MECH_FILE, REF_NAME, WORKLOAD
    shape, REF_NAME, WORKLOAD, states = WORKLOADS[name]
    if shape is not None:
        import tempfile
        from pyjac_b200 import synth
        MECH_FILE = os.path.join(tempfile.gettempdir(), 'pyjac_b200_bench_%s_%d.inp' % (shape, os.getpid()))
        synth.write(shape, MECH_FILE, seed=0)
    return states
METRIC = 'eval_jacob states/s'
UNIT = 'states/s'


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def bind_near_gpu(torch, dev):
    """Pinned staging buffers should live on the NUMA node the GPU hangs off: restrict this process to that node's
    CPUs before they are allocated (first touch places the pages).  Returns what was found and done, and the previous
    affinity for restore().  On hosts that expose a single node (the VMs of this pool) there is nothing to do."""
    info = {'nodes_visible': None, 'gpu_node': None, 'bound': False}
    prev = None
    try:
        nodes = sorted(int(d[4:]) for d in os.listdir('/sys/devices/system/node') if d.startswith('node') and d[4:].isdigit())
        info['nodes_visible'] = len(nodes)
        pr = torch.cuda.get_device_properties(dev)
        bdf = '%04x:%02x:%02x.0' % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open('/sys/bus/pci/devices/%s/numa_node' % bdf).read())
        info['gpu_node'] = node
        if len(nodes) > 1 and node >= 0:
            cpus = set()
            for part in open('/sys/devices/system/node/node%d/cpulist' % node).read().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus.update(range(int(lo), int(hi or lo) + 1))
            prev = os.sched_getaffinity(0)
            use = cpus & prev
            if use:
                os.sched_setaffinity(0, use)
                info['bound'] = True
                info['cpus'] = len(use)
    except Exception as exc:                              # sysfs layout / permissions: report, do not fail the bench
        info['note'] = str(exc).splitlines()[0][:120]
    return info, prev


def measured_peak():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as fh:
            return float(json.load(fh)['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


def ncu_traffic(nsp: int):
    """dram bytes per state from the committed ncu --set full capture (profiles/traffic.json)."""
    try:
        with open(os.path.join(ROOT, 'profiles', 'traffic.json')) as fh:
            return json.load(fh).get('%s_dram_bytes_per_state' % REF_NAME)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                 '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(',')]))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r for t, r in self.rows if t0 <= t <= t1 + 0.1 and len(r) >= 7] or \
               [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return None
        sm = []
        for r in rows:
            try:
                sm.append(float(r[0]))
            except ValueError:
                pass
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({nm for r in rows for nm, v in zip(names, r[3:7]) if v.lower().startswith('active')})
        try:
            smax = float(rows[0][1])
        except ValueError:
            smax = None
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': smax,
                'reasons': reasons, 'samples': len(rows)}


def load_states(nsp: int, n: int, seed: int):
    from pyjac_b200.states import synthetic_states
    return synthetic_states(nsp, n, seed=seed)


class Harness:
    """The reference's own `speedtest <num_odes> <num_threads>` (oracle/_ref/<name>/speedtest, built by
    oracle/build_ref.py from tester.c.in where the reference lies): reads data.bin, times eval_jacob over the
    states inside an OpenMP loop with its own timer.h, prints "num_odes,ms"."""

    def __init__(self, name, mech, P, y):
        import tempfile
        import numpy as np
        from pyjac_b200 import speedtest
        self.exe = os.path.join(ROOT, 'oracle', '_ref', name, 'speedtest')
        self.dir = tempfile.mkdtemp(prefix='pyjac_b200_bench_')
        Y_int = np.concatenate([y[:, 1:], 1.0 - y[:, 1:].sum(axis=1, keepdims=True)], axis=1)
        Y_orig = np.empty_like(Y_int)
        Y_orig[:, mech.fwd_spec_map] = Y_int            # data.bin holds the mechanism file's species order
        speedtest.write_data_bin(os.path.join(self.dir, 'data.bin'), y[:, 0], P, Y_orig)
        self.n_max = len(P)

    @staticmethod
    def available(name):
        return os.access(os.path.join(ROOT, 'oracle', '_ref', name, 'speedtest'), os.X_OK)

    def run(self, n, threads):
        """seconds the harness measured for n states"""
        out = subprocess.run([self.exe, str(n), str(threads)], cwd=self.dir, capture_output=True, text=True, check=True).stdout
        num, ms = out.strip().split(',')
        assert int(num) == n
        return float(ms) * 1e-3


def cpu_reference(mech, P, y):
    """The CPU baseline: the reference's generated C under its own harness when oracle/_ref holds it
    (kind 'reference'), else its generated C under oracle/ref_batch.c, else the oracle port.  Returns
    (run(n, threads) -> seconds, kind, description)."""
    from oracle.oracle import Oracle, RefLib
    if Harness.available(REF_NAME):
        h = Harness(REF_NAME, mech, P, y)
        return h.run, 'reference', "reference's generated C under its own speedtest harness (tester.c.in, timer.h), gcc -std=c99 -O3 -mtune=native -fopenmp"
    if RefLib.available(REF_NAME):
        ref = RefLib(REF_NAME)

        def run(n, threads):
            t = time.perf_counter()
            ref.eval_jacob(P[:n], y[:n], nthreads=threads, keep=False)
            return time.perf_counter() - t
        return run, 'reference', "reference's generated C (gcc -std=c99 -O3 -mtune=native -fopenmp) under an OpenMP loop over states"

    ora = Oracle(mech)

    def run(n, threads):
        t = time.perf_counter()
        ora.eval_jacob(P[:n], y[:n], nthreads=threads)
        return time.perf_counter() - t
    return run, 'port', 'oracle port (oracle/pyjac_oracle.c) on host cores'


def sample_size(run, threads, n_max, target_s):
    """number of states that keeps one pass near target_s seconds"""
    probe = min(n_max, 256 * threads)
    run(probe, threads)                         # warm caches / OpenMP pool
    rate = probe / run(probe, threads)
    return int(max(probe, min(n_max, rate * target_s)))


def run_reference(args, rank, world):
    if rank != 0:
        return
    from pyjac_b200.mechanism import Mechanism
    mech = Mechanism.from_chemkin(MECH_FILE)
    threads = host_threads()
    P, y = load_states(mech.NSP, 1 << 18, seed=0)
    run, kind, how = cpu_reference(mech, P, y)
    n = sample_size(run, threads, len(P), 2.0)            # bounded sample per step: ~2 s of CPU work
    for _ in range(args.warmup):
        run(n, threads)
    dt = sum(run(n, threads) for _ in range(args.steps))
    value = n * args.steps / dt
    sample = '%d states/step of the seed-0 synthetic batch, %d OpenMP threads' % (n, threads)
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': dt / args.steps * 1e3,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': WORKLOAD, 'states_per_step': n, 'path': how + ' on host cores'},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': threads, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def time_kernel(ev, torch, P, y, jac, steps, warmup, barrier, max_over_ranks):
    """ms per launch of eval_jacob on state-fastest device arrays (CUDA events, max over ranks), launches"""
    SF = dict(y_layout='state_fastest', jac_layout='state_fastest')
    for _ in range(warmup):
        ev.eval_jacob(P, y, jac, **SF)
    barrier()
    l0 = ev.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ev.eval_jacob(P, y, jac, **SF)
    e1.record()
    barrier()
    return max_over_ranks(e0.elapsed_time(e1)) / steps, e0.elapsed_time(e1) / steps, ev.launches - l0


def side_workload(name, torch, dev, rank, world, barrier, max_over_ranks, peak):
    """One of BASELINE.json's other configurations, device-resident, a few launches: sub-record of the line."""
    import numpy as np
    from pyjac_b200 import synth
    from pyjac_b200.evaluator import Evaluator
    from pyjac_b200.mechanism import Mechanism
    import tempfile
    if name == 'h2o2_pasr':
        # configs[0]: the bundled PaSR states (1020, not 1024: SURVEY.md finding 2) of the 10-species H2/O2 mechanism
        mech = Mechanism.from_chemkin(os.path.join(ROOT, 'tests', 'golden', 'h2o2_n2.inp'))
        g = np.load(os.path.join(ROOT, 'tests', 'golden', 'h2o2_pasr.npz'))
        P_h, y_h = g['P'], g['y']
        what = 'H2/O2 (10 sp / 28 rxn), the 1020 PaSR states of data/h2_pasr_output.npy'
    else:
        shape, _, what, n = SIDE[name]
        path = os.path.join(tempfile.gettempdir(), 'pyjac_b200_bench_%s_%d.inp' % (shape, os.getpid()))
        synth.write(shape, path, seed=0)
        mech = Mechanism.from_chemkin(path)
        P_h, y_h = load_states(mech.NSP, n, seed=rank)
    nsp, n = mech.NSP, len(P_h)
    ev = Evaluator(mech, dev.index)
    P = torch.tensor(P_h, device=dev)
    y = torch.tensor(y_h, device=dev).t().contiguous()
    jac = torch.empty((nsp * nsp, n), dtype=torch.float64, device=dev)
    ms, ms_local, launches = time_kernel(ev, torch, P, y, jac, 3, 2, barrier, max_over_ranks)
    bps = 8 * nsp * nsp + 8 * (nsp + 1)
    rec = {'workload': what, 'states_per_gpu': n, 'value': n * world / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms,
           'roofline_frac': bps * n / (ms_local * 1e-3) / 1e9 / peak, 'achieved_gbs': bps * n / (ms_local * 1e-3) / 1e9,
           'bytes_per_state': bps, 'kernel': ev.kernel_name(0),
           'plan': 'gs=%d, %d threads, working set in %s memory' % (ev.plan_gs, ev.plan_threads,
                                                                    'global' if int(ev.tables['p5_cfg'][14]) else 'shared')}
    ev.close()
    del jac, y, P
    torch.cuda.empty_cache()
    return rec


SIDE = {'usc2': WORKLOADS['usc2'], 'nc7': ('nc7', 'nc7', WORKLOADS['nc7'][2], 32768)}


def gather_run(ev, torch, dist, dev, rank, world, P, y, jac, nsp, n, steps, barrier, max_over_ranks):
    """eval_jacob of the rank's batch in chunks with every chunk's Jacobians gathered to rank 0 (NCCL
    send / recv on a second stream while the next chunk's kernel runs).  Rank 0 receives into a ring of
    two buffers per peer -- a consumer would work on a chunk and let it go: 8 x 23.6 GB do not fit one
    GPU (SURVEY.md finding 5)."""
    nchunk = next(c for c in (8, 4, 2, 1) if n % c == 0)
    cs = n // nchunk
    nn = nsp * nsp
    comm = torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    SF = dict(y_layout='state_fastest', jac_layout='state_fastest')
    # a chunk is a column block of the state-fastest arrays, evaluated into a buffer of its own leading
    # dimension, so that what is sent is contiguous; two buffers alternate
    out = [torch.empty((nn, cs), dtype=torch.float64, device=dev) for _ in range(2)]
    ring = [torch.empty((world - 1, nn, cs), dtype=torch.float64, device=dev) for _ in range(2)] if rank == 0 else None
    Pc = [P[c * cs:(c + 1) * cs].contiguous() for c in range(nchunk)]
    yc = [y[:, c * cs:(c + 1) * cs].contiguous() for c in range(nchunk)]

    def one_pass():
        sent = []
        for c in range(nchunk):
            if c >= 2:
                main.wait_event(sent[c - 2])           # the buffer is free once chunk c - 2 has gone
            ev.eval_jacob(Pc[c], yc[c], out[c & 1], **SF)
            done = torch.cuda.Event()
            done.record(main)
            with torch.cuda.stream(comm):
                comm.wait_event(done)
                if rank:
                    dist.send(out[c & 1], 0)
                else:
                    for q in [dist.irecv(ring[c & 1][r - 1], r) for r in range(1, world)]:
                        q.wait()
                e = torch.cuda.Event()
                e.record(comm)
            sent.append(e)
        main.wait_stream(comm)

    one_pass()                                     # warm-up: NCCL channels, staging
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one_pass()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1)) / steps
    return {'value': n * world / (ms * 1e-3), 'unit': UNIT, 'ms_per_step': ms, 'chunks': nchunk,
            'ingress_gbs': (world - 1) * n * nn * 8 / (ms * 1e-3) / 1e9,
            'how': 'NCCL send / recv of %d chunks per rank on a second stream, overlapped with the next chunk\'s '
                   'kernel; rank 0 receives into a ring of two buffers per peer' % nchunk}


def factored_runs(ev, torch, dev, mech, P_h, y_h, n_e, world, e_steps, barrier, max_over_ranks, j_chk):
    """Side records: (1) the host call that returns the factored record instead of the dense Jacobian
    (pyjac_eval_jacob_factored_host: 8 NF instead of 8 NSP^2 bytes per state over PCIe); (2) the consumer chain on
    the device -- pinned host states in, factored record, x = (I - gamma J)^-1 r, only x (NSP doubles) back."""
    nsp = mech.NSP
    nf, nnz = ev.factored_size
    yp = torch.empty((n_e, nsp), dtype=torch.float64, pin_memory=True)
    Pp = torch.empty((n_e,), dtype=torch.float64, pin_memory=True)
    fp = torch.empty((n_e, nf), dtype=torch.float64, pin_memory=True)
    yp.numpy()[:] = y_h[:n_e]
    Pp.numpy()[:] = P_h[:n_e]
    y_np, P_np, f_np = yp.numpy(), Pp.numpy(), fp.numpy()
    ev.eval_jacob_factored_host(P_np, y_np, f_np)                  # warm-up (staging buffers)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        ev.eval_jacob_factored_host(P_np, y_np, f_np)
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    k = len(j_chk)
    dense = ev.expand_factored(f_np[:k])
    err = np.abs(dense - j_chk) / (np.abs(j_chk).reshape(k, nsp, nsp).max(axis=2, keepdims=True).repeat(nsp, axis=2).reshape(k, -1) + 1e-300)
    assert err.max() <= 1e-13, 'factored record does not expand to the dense Jacobian (%.2e)' % err.max()
    fac_rec = {'value': n_e * world * e_steps / dt, 'unit': UNIT, 'h2d_bytes_per_step': n_e * (nsp + 1) * 8,
               'd2h_bytes_per_step': n_e * nf * 8, 'states_per_step': n_e, 'steps': e_steps,
               'doubles_per_state': nf, 'dense_doubles_per_state': nsp * nsp, 'sparse_block_entries': nnz,
               'api': 'pyjac_eval_jacob_factored_host (pinned host rows in, factored records out: energy row, T column, '
                      'rank-2 factors, sparse block in a fixed pattern); a sample is expanded on the host and compared '
                      'with the dense Jacobians'}
    del fp
    # consumer chain, chunked like the host API (2 streams would overlap copies with compute; kept simple: one stream)
    chunk = min(n_e, 1 << 18)
    fac = torch.empty((nf, chunk), dtype=torch.float64, device=dev)
    yd = torch.empty((chunk, nsp), dtype=torch.float64, device=dev)
    Pd = torch.empty((chunk,), dtype=torch.float64, device=dev)
    xd = torch.empty((chunk, nsp), dtype=torch.float64, device=dev)
    xp = torch.empty((n_e, nsp), dtype=torch.float64, pin_memory=True)
    rd = torch.ones((chunk, nsp), dtype=torch.float64, device=dev) * 1e-3
    gamma = 1.0e-6

    def chain():
        for s0 in range(0, n_e, chunk):
            cn = min(chunk, n_e - s0)
            yd[:cn].copy_(yp[s0:s0 + cn], non_blocking=True)
            Pd[:cn].copy_(Pp[s0:s0 + cn], non_blocking=True)
            # y arrives as rows; the record is written state-fastest (coalesced for the consumer kernel)
            ev.eval_jacob_factored(Pd[:cn], yd[:cn], out=fac, fac_layout='state_fastest')
            ev.newton_solve(fac, gamma, rd[:cn], out=xd[:cn], fac_layout='state_fastest')
            xp[s0:s0 + cn].copy_(xd[:cn], non_blocking=True)
        torch.cuda.synchronize()
    chain()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e_steps):
        chain()
    dt = max_over_ranks(time.perf_counter() - t0)
    # device-only times of the two kernels on one chunk
    e0, e1, e2_ = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    ev.eval_jacob_factored(Pd, yd, out=fac, fac_layout='state_fastest')
    e1.record()
    ev.newton_solve(fac, gamma, rd, out=xd, fac_layout='state_fastest')
    e2_.record()
    torch.cuda.synchronize()
    assert torch.isfinite(xd).all()
    cons = {'value': n_e * world * e_steps / dt, 'unit': UNIT, 'h2d_bytes_per_step': n_e * (nsp + 1) * 8,
            'd2h_bytes_per_step': n_e * nsp * 8, 'states_per_step': n_e, 'steps': e_steps,
            'factored_kernel_states_per_s': chunk / (e0.elapsed_time(e1) * 1e-3),
            'newton_kernel_states_per_s': chunk / (e1.elapsed_time(e2_) * 1e-3),
            'what': 'pinned host states in -> factored record (k_eval M_FACT) -> x = (I - gamma J)^-1 r per state (k_newton: '
                    'matrix expanded in shared memory, LU with partial pivoting) -> x back to pinned host memory; the dense '
                    'Jacobian is never written to HBM or sent over PCIe'}
    return fac_rec, cons


def run_ours(args, rank, world, local_rank):
    import numpy as np
    import torch
    import torch.distributed as dist
    from pyjac_b200.evaluator import Evaluator
    from pyjac_b200.mechanism import Mechanism

    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device -- pyjac_b200 has no CPU fallback')
    torch.cuda.set_device(local_rank)
    dev = torch.device('cuda', local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    mech = Mechanism.from_chemkin(MECH_FILE)
    nsp = mech.NSP
    n = args.states
    bytes_per_state = 8 * nsp * nsp + 8 * (nsp + 1)
    ev = Evaluator(mech, local_rank)
    wsg = bool(int(ev.tables['p5_cfg'][14]))
    # every rank gets its own shard of the (n * world)-state batch: seed = rank
    P_h, y_h = load_states(nsp, n, seed=rank)
    P = torch.tensor(P_h, device=dev)
    y = torch.tensor(y_h, device=dev).t().contiguous()              # [NSP][n], state-fastest
    jac = torch.empty((nsp * nsp, n), dtype=torch.float64, device=dev)
    SF = dict(y_layout='state_fastest', jac_layout='state_fastest')

    # ---- device-resident timing -------------------------------------------------------
    for _ in range(args.warmup):
        ev.eval_jacob(P, y, jac, **SF)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    barrier()
    if sampler:
        sampler.start()
        time.sleep(0.3)
    l0 = ev.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    w0 = time.time()
    e0.record()
    for _ in range(args.steps):
        ev.eval_jacob(P, y, jac, **SF)
    e1.record()
    barrier()
    w1 = time.time()
    launches = ev.launches - l0
    ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(w0, w1) if sampler else None
    ms_per_step = ms / args.steps
    value = n * world * args.steps / (ms * 1e-3)
    kernel_ms = e0.elapsed_time(e1) / max(launches, 1)         # this rank's average launch
    achieved = bytes_per_state * n / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()
    traffic = ncu_traffic(nsp)
    kernel_name = ev.kernel_name(0)

    # ---- the multi-GPU path of the north star: Jacobians gathered to rank 0; strong scaling ----
    with_gather = strong = None
    if world > 1 and not args.no_gather:
        with_gather = gather_run(ev, torch, dist, dev, rank, world, P, y, jac, nsp, n, max(1, min(args.steps, 3)),
                                 barrier, max_over_ranks)
    if world > 1:
        ns = (1 << 20) // world                               # BASELINE configs[1]'s batch, split over the ranks
        ms_s, _, _ = time_kernel(ev, torch, P[:ns].contiguous(), y[:, :ns].contiguous(),
                                 torch.empty((nsp * nsp, ns), dtype=torch.float64, device=dev), max(1, min(args.steps, 5)), 2,
                                 barrier, max_over_ranks)
        strong = {'global_states': ns * world, 'states_per_gpu': ns, 'value': ns * world / (ms_s * 1e-3), 'unit': UNIT,
                  'ms_per_step': ms_s}

    # ---- end to end through the host-pointer C-ABI call -------------------------------
    e2e = None
    if not args.no_e2e:
        import psutil
        numa, prev_aff = bind_near_gpu(torch, dev)
        n_e = n
        need = n_e * nsp * nsp * 8
        avail = psutil.virtual_memory().available // max(1, min(world, torch.cuda.device_count()))
        while n_e > 4096 and need * 1.5 > avail:
            n_e //= 2
            need = n_e * nsp * nsp * 8
        yp = torch.empty((n_e, nsp), dtype=torch.float64, pin_memory=True)
        Pp = torch.empty((n_e,), dtype=torch.float64, pin_memory=True)
        jp = torch.empty((n_e, nsp * nsp), dtype=torch.float64, pin_memory=True)
        yp.numpy()[:] = y_h[:n_e]
        Pp.numpy()[:] = P_h[:n_e]
        y_np, P_np, j_np = yp.numpy(), Pp.numpy(), jp.numpy()
        e_steps = max(1, min(args.steps, args.e2e_steps))
        ev.eval_jacob_host(P_np, y_np, j_np)                       # warm-up (staging buffers)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            ev.eval_jacob_host(P_np, y_np, j_np)                   # synchronous: returns after D2H
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        dt = max_over_ranks(dt)
        # spot-check that the host path delivered the same Jacobians as the device path
        if n_e == n:
            k = min(n, 64)
            assert np.array_equal(j_np[:k], jac[:, :k].t().cpu().numpy()), 'host / device API mismatch'
        j_chk = j_np[:64].copy()
        e2e = {'value': n_e * world * e_steps / dt, 'unit': UNIT,
               'h2d_bytes_per_step': n_e * (nsp + 1) * 8, 'd2h_bytes_per_step': n_e * nsp * nsp * 8,
               'states_per_step': n_e, 'steps': e_steps,
               'd2h_gbs': n_e * world * e_steps * nsp * nsp * 8 / dt / 1e9,
               'numa': numa,
               'api': 'pyjac_eval_jacob_host (pinned host rows in, pinned host Jacobians out)'}
        del yp, Pp, jp

    # ---- the same call with the factored record out, and the consumer step on the device (8 f1)
    e2e_factored = consumer = None
    if e2e is not None and not args.no_factored and world == 1:        # single-GPU side records (no collective inside: a
        # failure on one rank must not leave the others waiting)
        try:
            e2e_factored, consumer = factored_runs(ev, torch, dev, mech, P_h, y_h, e2e['states_per_step'], world,
                                                   e_steps, barrier, max_over_ranks, j_chk)
        except Exception as exc:                      # side records must not cost the headline line
            e2e_factored = {'error': str(exc).splitlines()[0][:200]}

    if not args.no_e2e and prev_aff is not None:
        os.sched_setaffinity(0, prev_aff)                 # the CPU baseline below uses every host thread again

    # ---- BASELINE.json's other configurations on this rank count ------------------------
    del jac
    torch.cuda.empty_cache()
    workloads = None
    if not args.no_workloads and args.workload == 'gri30':
        workloads = {}
        for name in ('usc2', 'nc7', 'h2o2_pasr'):
            try:
                workloads[name] = side_workload(name, torch, dev, rank, world, barrier, max_over_ranks, peak)
            except Exception as exc:                  # a side record must not cost the headline line
                workloads[name] = {'error': str(exc).splitlines()[0][:200]}

    # ---- CPU baseline (rank 0, single GPU runs only) ----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        threads = host_threads()
        ns_ = min(n, 1 << 18)
        run, kind, how = cpu_reference(mech, P_h[:ns_], y_h[:ns_])
        nc = sample_size(run, threads, ns_, args.cpu_seconds)
        dt = run(nc, threads)
        cpu = {'value': nc / dt, 'unit': UNIT, 'cores': threads, 'kind': kind,
               'sample': 'first %d states of the same batch, %.1f s, %d OpenMP threads; %s' % (nc, dt, threads, how)}

    if rank == 0:
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': {'workload': WORKLOAD, 'states_per_gpu': n, 'global_states': n * world,
                       'jacobian_bytes_per_gpu': n * nsp * nsp * 8,
                       'l2': 'inputs (%.0f MB) and outputs (%.1f GB) per step exceed the 126 MB L2'
                             % (n * (nsp + 1) * 8 / 1e6, n * nsp * nsp * 8 / 1e9),
                       'layout': 'state-fastest (struct-of-arrays) in and out: y[NSP][n], jac[NSP*NSP][n] -- the '
                                 "reference's GPU layout; e2e: one row per state in, one column-major "
                                 'NSPxNSP Jacobian per state out (the scalar API layout)',
                       'plan': 'gs=%d states per block, %d threads, working set in %s memory'
                               % (ev.plan_gs, ev.plan_threads, 'global' if wsg else 'shared'),
                       'parallelism': 'state batch sharded over %d GPU(s), no collective in the timed region of '
                                      '`value`; `with_gather` adds the NCCL gather of the Jacobians to rank 0' % world},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': None if traffic is None else traffic * n,
                         'traffic_source': None if traffic is None else
                         'profiles/traffic.json: dram__bytes_read.sum + dram__bytes_write.sum per state of an ncu --set full '
                         'capture of this kernel (65 536 states), scaled to this batch -- not measured in this run',
                         'peak_source': peak_src, 'bytes_per_state': bytes_per_state,
                         'kernel': kernel_name, 'kernel_ms': kernel_ms},
            'e2e': e2e, 'e2e_factored': e2e_factored, 'consumer': consumer, 'with_gather': with_gather, 'strong': strong, 'workloads': workloads,
            'cpu_baseline': cpu, 'gpu_launches': launches, 'clocks': clocks,
        }
        print(json.dumps(line), flush=True)
    ev.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='gri30', choices=sorted(WORKLOADS))
    ap.add_argument('--states', type=int, default=0, help='states per GPU per step (0 = the workload default)')
    ap.add_argument('--e2e-steps', type=int, default=3)
    ap.add_argument('--cpu-seconds', type=float, default=10.0)
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-factored', action='store_true', help='skip the factored-record / consumer side records')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-workloads', action='store_true', help='leave out the sub-records of the other BASELINE configurations')
    ap.add_argument('--no-gather', action='store_true', help='(N > 1) leave out the gather-to-rank-0 measurement')
    args = ap.parse_args()
    default_states = select_workload(args.workload)
    args.states = args.states or default_states
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if world == 1 and args.gpus > 1 and args.impl == 'ours':
        # not launched by torchrun: re-exec under it so that one rank drives each GPU
        cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
               '--nproc-per-node', str(args.gpus), '--master-addr', '127.0.0.1',
               '--master-port', str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == 'reference':
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == '__main__':
    main()

"""Developer script: the tables under profiles/ from `ncu --set full --import-source on` reports.

    python scripts/ncu_summary.py summary out.csv a.ncu-rep [b.ncu-rep ...]
        one row per captured kernel launch with the metrics DESIGN.md / profiles/README.md quote
    python scripts/ncu_summary.py lines out.csv a.ncu-rep k_helium [n_models_x_steps]
        the source lines of one kernel with the most executed thread instructions, their share of the
        stall samples, and the size of the hot loop (SASS instructions executed at least once per
        ten model-steps)
    python scripts/ncu_summary.py sass out.txt a.ncu-rep k_helium n_model_steps
        the hot loop of one kernel as SASS, in address order: every instruction executed at least once
        per ten model-steps with its warp-level and thread-level execution counts (lanes per issue)
Needs `ncu` on PATH (reads reports; no GPU)."""
import collections
import csv
import io
import subprocess
import sys

METRICS = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread',
           'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
           'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
           'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
           'smsp__average_warp_latency_per_inst_issued.ratio',
           'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
           'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
           'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sass__inst_executed_local_loads',
           'sass__inst_executed_local_stores', 'l1tex__t_sector_hit_rate.pct']


def ncu(args):
    return subprocess.run(['ncu'] + args, check=True, capture_output=True, text=True).stdout


def summary(out, reports):
    rows_out = [['kernel'] + METRICS]
    for rep in reports:
        rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'raw', '--csv']))))
        hdr, units = rows[0], rows[1]
        ik = hdr.index('Kernel Name')
        for r in rows[2:]:
            if len(r) < len(hdr):
                continue
            line = [r[ik].split('(')[0]]
            for m in METRICS:
                line.append('%s %s' % (r[hdr.index(m)], units[hdr.index(m)]) if m in hdr else '')
            rows_out.append(line)
    with open(out, 'w', newline='') as fh:
        csv.writer(fh).writerows(rows_out)


def lines(out, rep, kernel, model_steps, top=40):
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass',
                                            '-k', 'regex:' + kernel]))))
    cur, hdr, line = None, None, None
    thr, smp, src = collections.Counter(), collections.Counter(), {}
    hot, static = 0, 0
    seen = set()
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if len(r) == 2:
            continue
        if r and r[0] == 'Line No':
            hdr = r
            i_i, i_t, i_s = (hdr.index(k) for k in ('Instructions Executed', 'Thread Instructions Executed', '# Samples'))
            continue
        if hdr is None or len(r) <= i_t:
            continue
        if r[0] != '':
            line = (cur, int(r[0]))
            src[line] = r[1].strip()
            continue
        if not r[2].startswith('0x') or r[2] in seen:   # (a full-set report lists the SASS once per view)
            continue
        seen.add(r[2])
        try:
            n_t, n_s, n_i = int(r[i_t]), int(r[i_s]), int(r[i_i])
        except ValueError:
            continue
        static += 1
        thr[line] += n_t
        smp[line] += n_s
        if model_steps and n_t >= model_steps / 10:
            hot += 1
    tt, ts = sum(thr.values()), sum(smp.values())
    with open(out, 'w', newline='') as fh:
        w = csv.writer(fh)
        w.writerow(['# %s: %d SASS instructions, %d executed at least once per ten model-steps (%.1f KB); '
                    '%.0f thread instructions per model-step' % (kernel, static, hot, hot * 16 / 1024,
                                                                tt / model_steps if model_steps else 0)])
        w.writerow(['file', 'line', 'pct_of_thread_instructions', 'pct_of_stall_samples', 'source'])
        for k, v in thr.most_common(top):
            w.writerow([k[0], k[1], '%.2f' % (100 * v / tt), '%.2f' % (100 * smp[k] / ts), src[k][:100]])


def sass(out, rep, kernel, model_steps):
    rows = list(csv.reader(io.StringIO(ncu(['-i', rep, '--page', 'source', '--csv', '--print-source', 'sass',
                                            '-k', 'regex:' + kernel]))))
    hdr, body = None, []
    for r in rows:
        if r and r[0] == 'Address':
            hdr = r
            continue
        if hdr is not None and len(r) >= 8 and r[0].startswith('0x'):
            body.append(r)
    i_a = hdr.index('Address') if 'Address' in hdr else 0
    i_src = hdr.index('Source')
    i_i, i_t, i_s = (hdr.index(k) for k in ('Instructions Executed', 'Thread Instructions Executed', '# Samples'))
    seen, keep, tot_t, tot_i = set(), [], 0, 0
    for r in body:
        if r[i_a] in seen:
            continue
        seen.add(r[i_a])
        try:
            n_i, n_t, n_s = int(r[i_i]), int(r[i_t]), int(r[i_s])
        except ValueError:
            continue
        tot_t += n_t
        tot_i += n_i
        if n_t >= model_steps / 10:
            keep.append((r[i_a], n_i, n_t, n_s, r[i_src]))
    with open(out, 'w') as fh:
        fh.write('# %s: %d SASS instructions, %d in the hot loop (executed at least once per ten model-steps, '
                 '%.1f KB); %.0f thread instructions and %.0f warp instructions per model-step\n'
                 % (kernel, len(seen), len(keep), len(keep) * 16 / 1024, tot_t / model_steps, tot_i / model_steps))
        fh.write('# address  warp_instructions  thread_instructions  lanes_per_issue  stall_samples  sass\n')
        for a, n_i, n_t, n_s, src in keep:
            fh.write('%s %11d %12d %5.2f %6d  %s\n' % (a, n_i, n_t, n_t / max(n_i, 1), n_s, src))


if __name__ == '__main__':
    if sys.argv[1] == 'summary':
        summary(sys.argv[2], sys.argv[3:])
    elif sys.argv[1] == 'sass':
        sass(sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5]))
    else:
        lines(sys.argv[2], sys.argv[3], sys.argv[4], float(sys.argv[5]) if len(sys.argv) > 5 else 0)

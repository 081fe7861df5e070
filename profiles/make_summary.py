"""Writes profiles/r2_summary.md from the committed records (bench lines, scaling runs, ncu summaries):
    python profiles/make_summary.py"""
import json
import os

ROOT = os.path.dirname(os.path.abspath(__file__)) + '/'


def line(path):
    return json.loads(open(ROOT + path).read().strip().splitlines()[-1])


def main():
    b, r, k100 = line('bench_r2_1gpu.json'), line('bench_r2_reference.json'), line('bench_r2_1gpu_top100.json')
    scaling = [('r3g_n2.json', '2 GPUs, query groups (default)'), ('r3g_n4.json', '4 GPUs, query groups (default)'),
               ('r3g_n8.json', '8 GPUs, query groups (default)'), ('r3g_n8t2.json', '8 GPUs, 2 truth shards x 4 query groups'),
               ('r3g_n8t8.json', '8 GPUs, 8 truth shards (shared thresholds)'),
               ('r2p_n8t8_noshare.json', '8 GPUs, 8 truth shards, local thresholds only (x2 sweep)'),
               ('r2p_n4t4.json', '4 GPUs, 4 truth shards (x2 sweep)'), ('r2l_n2t2_shared.json', '2 GPUs, 2 truth shards (x2 sweep)'),
               ('r2p_c5n8.json', 'C5: 1M x 10M, 8 GPUs, 8 truth shards (x2 sweep)')]
    t1 = b['ms_per_step']
    rf, e = b['roofline'], b['extra']
    c = e['c1_c2']
    out = ['# Round 2 - measured summary (B200, sm_100a; every figure below has its raw file in this directory)\n',
           '## Headline: C3 = 100,000 test x 500,000 truth titles, top-10, one GPU (`bench_r2_1gpu.json`, driver command line)\n',
           f"* device path **{b['ms_per_step']:.2f} ms per step = {b['value'] / 1e6:.2f} M titles/s**; end to end from pinned host buffers "
           f"{b['e2e']['value'] / 1e6:.2f} M titles/s ({b['e2e']['h2d_bytes_per_step'] / 1e6:.1f} MB H2D + {b['e2e']['d2h_bytes_per_step'] / 1e6:.1f} MB D2H inside "
           f"the timed region); {b['gpu_launches'] // b['steps']} kernel launches per step",
           f"* workload: top trigram in {b['config']['top_trigram_df_share'] * 100:.1f} % of the titles, **{b['config']['postings_hit_per_query_over_n']:.2f} N "
           f"postings per query** (example data 0.89 N, round 1's generator 0.44 N), {b['config']['mean_trigrams_per_truth_title']:.1f} trigrams per title",
           f"* parity: {b['parity']['checked_queries']} sampled queries vs the CPU oracle: {b['parity']['mismatching_queries']} mismatches; "
           f"{b['parity']['reference_numba_checked_queries']} queries vs the staged reference's own numba kernels: "
           f"{b['parity']['reference_numba_mismatching_queries']} mismatches",
           f"* CPU: the reference's numba kernels {r['value']:.0f} titles/s on {r['cpu_baseline']['cores']} host cores (`bench_r2_reference.json`, "
           f"{r['parity']['reference_vs_port_checked_queries']} queries per step, {r['parity']['reference_vs_port_mismatching_queries']} mismatches against the C "
           f"port); the C / OpenMP port {b['cpu_baseline']['port']['value']:.0f} titles/s",
           f"* top_n = 100: {k100['ms_per_step']:.1f} ms per step (`bench_r2_1gpu_top100.json`); real example titles tiled to the same size: "
           f"{e['c3_example']['ms_per_step']:.1f} ms ({e['c3_example']['statistics']['postings_hit_per_query_over_n']:.2f} N postings per query, "
           f"{e['c3_example']['parity']['mismatching_queries']} mismatches of {e['c3_example']['parity']['checked_queries']})\n",
           '## Dominant kernel `k_post` against its measured limiter (`r2_k_post_ncu.json`, `r2_k_post_metrics.csv`)\n',
           f"* share of the step {rf['kernel_share_of_step'] * 100:.0f} % ({rf['k1_kernels']['k_post']['ms_per_step']:.1f} ms in "
           f"{rf['k1_kernels']['k_post']['launches_per_step']:.0f} launches, measured live with CUDA events); ncu launch list `r2_launches.md`: 85 % (cold cache, "
           f"serialised)",
           f"* bound: **{rf['bound']}**, utilisation {rf['frac']:.2f} (issue slots {rf['ncu']['issue_active_pct']:.0f} %, shared-memory wavefronts "
           f"{rf['ncu']['smem_wavefronts_pct']:.0f} %, DRAM throughput {rf['ncu']['dram_throughput_pct']:.2f} %, warp slots occupied "
           f"{rf['ncu']['warps_active_pct']:.0f} %, {rf['ncu']['registers']:.0f} registers); DRAM traffic {rf['traffic'] / 1e6:.1f} MB per launch; ncu taken from "
           f"the current kernel sources: {rf['ncu_matches_current_sources']}",
           f"* SURVEY 8(d) byte convention (labelled secondary): {rf['hbm_convention']['achieved_gbs'] / 1e3:.1f} TB/s = {rf['hbm_convention']['frac']:.1f} x the "
           f"measured HBM peak - the index is L2 resident, the kernel does not move those bytes\n",
           '## Other kernels (ncu summaries `r2_<kernel>_ncu.json`)\n',
           '| kernel | workload | issue slots | shared-memory wavefronts | DRAM throughput | captured launch |', '|---|---|---:|---:|---:|---:|']
    for k in ('k_select', 'k_scan', 'k_indel_groups', 'k_indel_pairs', 'k_feature_words'):
        j = json.load(open(ROOT + f'r2_{k}_ncu.json'))
        out.append(f"| `{k}` | {'C3 step' if k in ('k_select', 'k_scan') else '200k candidate pairs'} | {j['issue_active_pct']:.0f} % | "
                   f"{j['smem_wavefronts_pct']:.0f} % | {j['dram_throughput_pct']:.2f} % | {j['captured_launch_ms'] * 1e3:.0f} us |")
    out += ['', '## Pair kernels and the other BASELINE configs (all inside `bench_r2_1gpu.json` -> `extra`)\n',
            f"* candidate pairs of the step (1M): InDel ratio {e['candidate_pairs']['indel_ratio']['pairs_per_s'] / 1e9:.2f} G pairs/s, construct_features "
            f"{e['candidate_pairs']['construct_features']['pairs_per_s'] / 1e6:.0f} M pairs/s; title strings -> index build -> candidates -> 1M x 66 features: "
            f"{e['candidate_pairs']['pipeline']['ms']:.1f} ms",
            f"* C4 (100M pairs, 10 % of the titles 65..128 characters): InDel ratio {e['c4']['indel_ratio']['pairs_per_s'] / 1e9:.2f} G pairs/s, construct_features "
            f"{e['c4']['construct_features']['pairs_per_s'] / 1e6:.0f} M pairs/s; C port on {e['c4']['cpu_baseline']['cores']} cores "
            f"{e['c4']['cpu_baseline']['indel_ratio_pairs_per_s'] / 1e6:.1f} M / {e['c4']['cpu_baseline']['construct_features_pairs_per_s'] / 1e6:.2f} M; parity sample "
            f"of {e['c4']['parity']['sampled_pairs']} pairs: {e['c4']['parity']['ratio_mismatches']} ratio, {e['c4']['parity']['integer_feature_mismatches']} integer-feature, "
            f"{e['c4']['parity']['float_feature_mismatches']} float-feature mismatches",
            f"* C1 (example data through the reference's API, top_n = 100): index build {c['c1_candidates']['ours']['index_build_s']:.2f} s vs "
            f"{c['c1_candidates']['reference']['index_build_s']:.1f} s; {c['c1_candidates']['ours']['titles_per_s']:.0f} vs "
            f"{c['c1_candidates']['reference']['titles_per_s']:.0f} titles/s incl. the build; pre-match {c['c1_prematch']['ours']['pairs_per_s'] / 1e6:.1f} M vs "
            f"{c['c1_prematch']['reference']['pairs_per_s'] / 1e6:.2f} M pairs/s; features {c['c1_features']['ours']['pairs_per_s'] / 1e6:.1f} M vs "
            f"{c['c1_features']['reference']['pairs_per_s'] / 1e6:.2f} M pairs/s; mismatches: {c['c1_candidates']['parity']['mismatching_rows']} rows, "
            f"{c['c1_prematch']['parity']['mismatching_pairs']} ratios, {c['c1_features']['parity']['integer_feature_mismatches']} + "
            f"{c['c1_features']['parity']['float_feature_mismatches']} features",
            f"* C2 (single title vs 30,000 truth titles, index build included): {c['c2_single_title']['ours']['latency_s'] * 1e3:.0f} ms "
            f"({c['c2_single_title']['ours']['canonical_order_latency_s'] * 1e3:.0f} ms canonical order) vs {c['c2_single_title']['reference']['latency_s']:.2f} s; same "
            f"candidates: {c['c2_single_title']['parity']['same_candidates']}\n",
            '## Multi-GPU (`r2_scaling/`, one 8 x B200 box, `--steps 20 --warmup 5`, device-timed max over ranks)\n',
            '| layout | ms per step | titles/s | e2e titles/s | efficiency t1 / (N t_N) | oracle sample: checked / mismatching |', '|---|---:|---:|---:|---:|---|']
    for name, label in scaling:
        s = line('r2_scaling/' + name)
        eff = f"{t1 / (s['n_gpus'] * s['ms_per_step']):.2f}" if 'C5' not in label else '-'
        out.append(f"| {label} | {s['ms_per_step']:.2f} | {s['value'] / 1e6:.2f} M | {s['e2e']['value'] / 1e6:.2f} M | {eff} | "
                   f"{s['parity']['checked_queries']} / {s['parity']['mismatching_queries']} |")
    out.append('\nC5 scores 1.0e13 (query, truth) pairs per step: 1.26e13 pairs/s on 8 GPUs (round 1, dense kernel: 2.04 s per step).')
    open(ROOT + 'r2_summary.md', 'w').write('\n'.join(out) + '\n')


if __name__ == '__main__':
    main()

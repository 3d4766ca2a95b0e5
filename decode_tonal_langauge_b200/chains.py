"""The two benchmark chains in the reference's YAML ``steps`` vocabulary.

EX   = ref: example_config.yaml:27-45 (the shipped example, CPU-runnable case).
FULL6 = the north-star chain: notch -> CAR -> band-pass -> Gaussian-Hilbert
envelope -> FFT downsample -> per-channel z-score.
"""

EX_STEPS = [
    {"module": "preprocess.downsample", "params": {"downsample_freq": 400}},
    {"module": "preprocess.frequency_filter", "params": {"bands": [
        {"method": "hilbert", "params": {"freq_ranges": [70.0, 150.0], "envelope": True}},
        {"method": "butter", "params": {"freqs": [0.3, 100], "filter_type": "bandpass"}},
    ]}},
    {"module": "preprocess.zscore_rereference", "params": {"rereference_interval": [0.0, 25.0]}},
]

FULL6_STEPS = [
    {"module": "preprocess.frequency_filter", "params": {"bands": [
        {"method": "butter", "params": {"freqs": [58, 62], "filter_type": "bandstop"}}]}},
    {"module": "preprocess.car_rereference", "params": {}},
    {"module": "preprocess.frequency_filter", "params": {"bands": [
        {"method": "butter", "params": {"freqs": [70, 150], "filter_type": "bandpass"}}]}},
    {"module": "preprocess.frequency_filter", "params": {"bands": [
        {"method": "hilbert", "params": {"freq_ranges": [70.0, 150.0], "envelope": True}}]}},
    {"module": "preprocess.downsample", "params": {"downsample_freq": 400}},
    {"module": "preprocess.channel_zscore", "params": {}},
]

CHAINS = {"EX": EX_STEPS, "FULL6": FULL6_STEPS}

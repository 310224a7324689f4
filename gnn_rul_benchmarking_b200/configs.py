"""FC_STGNN hyper-parameter sets of the reference (configs/hparams.py) that the bench and tests
name, plus the BASELINE.json synthetic shape.  Keys are exactly the `alg_hparams['FC_STGNN']`
kwargs passed to FC_STGNN_RUL(**configs) at trainer.py:61 / algorithms.py:58."""

CONFIGS = {
    # configs/hparams.py:149-151 (CMAPSS FD004) -- the metric config "S1"
    "FD004": dict(patch_size=2, num_patch=25, encoder_time_out=4, encoder_hidden_dim=8, encoder_out_dim=6,
                  encoder_conv_kernel=2, hidden_dim=8, num_sequential=10, num_node=14, num_windows=36),
    # configs/hparams.py:32-34 (FD001)
    "FD001": dict(patch_size=25, num_patch=2, encoder_time_out=27, encoder_hidden_dim=8, encoder_out_dim=32,
                  encoder_conv_kernel=2, hidden_dim=8, num_sequential=6, num_node=14, num_windows=2),
    # configs/hparams.py:69-71 (FD002)
    "FD002": dict(patch_size=1, num_patch=50, encoder_time_out=3, encoder_hidden_dim=8, encoder_out_dim=12,
                  encoder_conv_kernel=2, hidden_dim=8, num_sequential=10, num_node=14, num_windows=74),
    # configs/hparams.py:109-111 (FD003)
    "FD003": dict(patch_size=1, num_patch=50, encoder_time_out=3, encoder_hidden_dim=8, encoder_out_dim=6,
                  encoder_conv_kernel=2, hidden_dim=24, num_sequential=25, num_node=14, num_windows=74),
    # configs/hparams.py:196-198 (N-CMAPSS)
    "NCMAPSS": dict(patch_size=2, num_patch=25, encoder_time_out=4, encoder_hidden_dim=8, encoder_out_dim=32,
                    encoder_conv_kernel=2, hidden_dim=8, num_sequential=6, num_node=20, num_windows=36),
    # SURVEY.md section 8d "S2": BASELINE north_star synthetic [B,T=50,N=21,C=14]
    "S2": dict(patch_size=1, num_patch=50, encoder_time_out=3, encoder_hidden_dim=8, encoder_out_dim=6,
               encoder_conv_kernel=2, hidden_dim=7, num_sequential=10, num_node=21, num_windows=74),
}
# configs/hparams.py:133 train_params['FC_STGNN'] (identical for FD001-FD004 and N-CMAPSS)
TRAIN_PARAMS = dict(num_epochs=81, batch_size=100, weight_decay=1e-4, learning_rate=1e-3)

# ASTGCNN (BASELINE.json configs[2]): configs/hparams.py:38 (C-MAPSS, 14 sensors) and :202 (N-CMAPSS, 20 channels)
ASTGCNN_CONFIGS = {
    "CMAPSS": dict(num_nodes=14, time_length=50, encoder_out_dim=50, output_dim=64, K=3),
    "NCMAPSS": dict(num_nodes=20, time_length=50, encoder_out_dim=50, output_dim=64, K=3),
}

"""Default hyper-parameters of the distillation path (values of the reference's config.py:21-48).
Only the constants the hot path reads are carried; dataset / checkpoint paths are out of scope."""

# STFT front end (config.py:22-29)
fs = 16000
win_len = 400
win_inc = 100
fft_len = 512
window_type = 'hamming'

# teacher (config.py:31-37)
rnn_layers = 2
rnn_units = 256
masking_mode = 'E'
use_clstm = True
kernel_num = [32, 64, 128, 256, 256, 256]
loss_mode = 'SI-SNR'   # the reference default 'SDR+PMSQE' needs asteroid's PMSQE (out of scope)

# optimisation (config.py:40-42, distill.py:202-204)
max_epochs = 20
learning_rate = 0.0006
batch = 32
weight_decay = 5e-4

# student of the reference: quarter width (config.py:45-48)
rnn_layers_student = 2
rnn_units_student = 64
kernel_num_student = [8, 16, 32, 64, 64, 64]

# half-width student named by the benchmark configuration
rnn_units_half = 128
kernel_num_half = [16, 32, 64, 128, 128, 128]

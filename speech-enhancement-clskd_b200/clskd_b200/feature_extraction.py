"""Feature taps (reference: feature_extraction.py:3-50): forward hooks on the encoder / decoder
blocks and the complex LSTM of a DCCRN.  Works on the reference's model and on clskd_b200.DCCRN
alike, because the B200 model keeps the same submodule tree."""


class DCCRN:
    def __init__(self, model):
        self.model = model
        try:        # clskd_b200 models: sum the gradients of tapped tensors in the library
            from .ops import fanout
            self._fanout = fanout if type(model).__module__.startswith("clskd_b200") else None
        except Exception:      # used on a reference model
            self._fanout = None
        self.feature_maps = {"encoder": [], "decoder": [], "clstm": []}
        self._handles = []
        for blk in model.encoder:
            self._handles.append(blk.register_forward_hook(self._tap("encoder")))
        for blk in model.decoder:
            self._handles.append(blk.register_forward_hook(self._tap("decoder")))
        self._handles.append(model.enhance.register_forward_hook(self._tap("clstm")))
        # reference attribute names
        n_enc, n_dec = len(model.encoder), len(model.decoder)
        self.handle_encoder = self._handles[:n_enc]
        self.handle_decoder = self._handles[n_enc:n_enc + n_dec]
        self.handle_clstm = self._handles[-1]

    def _tap(self, key):
        def hook(module, inputs, output):
            # a tapped tensor has two consumers (the distillation loss and the rest of the model): hand out two
            # aliases whose gradients the library sums with one kernel (ops.fanout; identity outside autograd)
            fan = getattr(self, "_fanout", None)
            if fan is None:
                self.feature_maps[key].append(output)
                return None
            if isinstance(output, (list, tuple)):
                pairs = [fan(o, 2) if hasattr(o, "requires_grad") else (o, o) for o in output]
                self.feature_maps[key].append(type(output)(p[0] for p in pairs))
                return type(output)(p[1] for p in pairs)
            a, b = fan(output, 2)
            self.feature_maps[key].append(a)
            return b
        return hook

    # reference method names
    def encoder_hook(self, module, input, output):
        self.feature_maps["encoder"].append(output)

    def decoder_hook(self, module, input, output):
        self.feature_maps["decoder"].append(output)

    def enhance_hook(self, module, input, output):
        self.feature_maps["clstm"].append(output)

    def remove_hook(self):
        for h in self._handles:
            h.remove()
        self._handles = []

    def extract_feature_maps(self, input):
        self.model(input)
        return self.feature_maps

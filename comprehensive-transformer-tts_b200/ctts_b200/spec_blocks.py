"""Parameter inventories of the `transformer`, `fastformer` and `conformer` block types.

Names / shapes follow model/transformers/transformer.py:15-288, fastformer.py:16-376 and conformer.py:20-560
(probed against the reference's own state_dict: tests/test_state_dict_contract.py).  An init of "tie:<name>" means
the entry is THE SAME Parameter object as <name> (fastformer ties the attention-logit projections of every layer to
layer 0, fastformer.py:157-161; conformer re-registers the encoder's / decoder's positional table inside every
block, conformer.py:321).
"""
from .spec import _conv, _lin, _ln, vocab_size


def _abs_table(out, name, rows, d):
    out.append((name, (1, rows, d), "frozen", "sinusoid_interleaved"))


def _transformer(out, cfg):
    c = cfg["transformer"]
    n_pos = cfg["max_seq_len"] + 1
    k = c["conv_kernel_size"]
    f = c["conv_filter_size"]

    def stack(pre, n_layers, d):
        for i in range(n_layers):
            p = "%slayer_stack.%d." % (pre, i)
            for w in ("w_qs", "w_ks", "w_vs"):
                out.append((p + "slf_attn.%s.linear.weight" % w, (d, d), "param", "xavier"))
            _ln(out, p + "slf_attn.layer_norm", d)
            out.append((p + "slf_attn.fc.linear.weight", (d, d), "param", "xavier"))
            _conv(out, p + "pos_ffn.w_1", f, d, k[0])
            _conv(out, p + "pos_ffn.w_2", d, f, k[1])
            _ln(out, p + "pos_ffn.layer_norm", d)

    d_enc, d_dec = c["encoder_hidden"], c["decoder_hidden"]
    _abs_table(out, "encoder.position_enc", n_pos, d_enc)
    out.append(("encoder.src_word_emb.weight", (vocab_size(), d_enc), "param", "emb1"))
    stack("encoder.", c["encoder_layer"], d_enc)
    _abs_table(out, "decoder.position_enc", n_pos, d_dec)
    stack("decoder.", c["decoder_layer"], d_dec)
    return d_enc, d_dec


def _fastformer(out, cfg):
    c = cfg["transformer"]  # sic: fastformer.py:24-34 reads the `transformer` section
    n_pos = cfg["max_seq_len"] + 1
    k = c["conv_kernel_size"]
    f = c["conv_filter_size"]

    def stack(pre, n_layers, d, n_head):
        heads = d // n_head          # FastAttention(d_model, d_head, n_head): dim_head is used as the head COUNT
        for i in range(n_layers):
            a = "%slayer_stack.layers.%d.0." % (pre, i)
            g = "%slayer_stack.layers.%d.1." % (pre, i)
            first = "%slayer_stack.layers.0.0." % pre
            _ln(out, a + "norm", d)
            _lin(out, a + "fn.query", d, d, init="normal002")
            for nm in ("to_q_attn_logits",):
                if i == 0:
                    _lin(out, a + "fn." + nm, heads, d, init="normal002")
                else:
                    out.append((a + "fn.%s.weight" % nm, (heads, d), "param", "tie:" + first + "fn.%s.weight" % nm))
                    out.append((a + "fn.%s.bias" % nm, (heads,), "param", "tie:" + first + "fn.%s.bias" % nm))
            _lin(out, a + "fn.key", d, d, init="normal002")
            for nm in ("to_k_attn_logits",):
                if i == 0:
                    _lin(out, a + "fn." + nm, heads, d, init="normal002")
                else:
                    out.append((a + "fn.%s.weight" % nm, (heads, d), "param", "tie:" + first + "fn.%s.weight" % nm))
                    out.append((a + "fn.%s.bias" % nm, (heads,), "param", "tie:" + first + "fn.%s.bias" % nm))
            _lin(out, a + "fn.transform", d, d, init="normal002")
            _ln(out, g + "norm", d)
            _conv(out, g + "fn.w_1", f, d, k[0])
            _conv(out, g + "fn.w_2", d, f, k[1])

    d_enc, d_dec = c["encoder_hidden"], c["decoder_hidden"]
    _abs_table(out, "encoder.position_enc", n_pos, d_enc)
    out.append(("encoder.src_word_emb.weight", (vocab_size(), d_enc), "param", "emb1"))
    stack("encoder.", c["encoder_layer"], d_enc, c["encoder_head"])
    _abs_table(out, "decoder.position_enc", n_pos, d_dec)
    stack("decoder.", c["decoder_layer"], d_dec, c["decoder_head"])
    return d_enc, d_dec


def _conformer(out, cfg):
    c = cfg["conformer"]
    n_pos = cfg["max_seq_len"] + 1
    ff = c["feed_forward_expansion_factor"]
    assert c["conv_expansion_factor"] == 2, "Currently, Only Supports expansion_factor 2"  # conformer.py:457
    k = c["conv_kernel_size"]

    def stack(pre, n_layers, d, n_head):
        for i in range(n_layers):
            p = "%slayer_stack.%d.sequential." % (pre, i)
            for j in (0, 3):
                q = "%s%d.module.sequential." % (p, j)
                _ln(out, q + "0", d)
                _lin(out, q + "1.linear", d * ff, d, init="xavier")
                _lin(out, q + "4.linear", d, d * ff, init="xavier")
                if j == 3:
                    continue
                a = p + "1.module."
                out.append((a + "positional_encoding", (1, n_pos, d), "frozen", "tie:" + pre + "position_enc"))
                _ln(out, a + "layer_norm", d)
                out.append((a + "attention.u_bias", (n_head, d // n_head), "param", "xavier"))
                out.append((a + "attention.v_bias", (n_head, d // n_head), "param", "xavier"))
                for nm in ("query_proj", "key_proj", "value_proj", "pos_proj", "out_proj"):
                    out.append((a + "attention.%s.linear.weight" % nm, (d, d), "param", "xavier"))
                m = p + "2.module.sequential."
                _ln(out, m + "0", d)
                _conv(out, m + "2.conv", 2 * d, d, 1)
                out.append((m + "4.conv.weight", (d, 1, k), "param", "default:%d" % k))
                _ln(out, m + "5", d)
                out.append((m + "5.running_mean", (d,), "buffer", "zeros"))
                out.append((m + "5.running_var", (d,), "buffer", "ones"))
                out.append((m + "5.num_batches_tracked", (), "buffer", "count"))
                _conv(out, m + "7.conv", d, d, 1)
            _ln(out, p + "4", d)

    d_enc, d_dec = c["encoder_hidden"], c["decoder_hidden"]
    _abs_table(out, "encoder.position_enc", n_pos, d_enc)
    out.append(("encoder.src_word_emb.weight", (vocab_size(), d_enc), "param", "emb1"))
    stack("encoder.", c["encoder_layer"], d_enc, c["encoder_head"])
    _abs_table(out, "decoder.position_enc", n_pos, d_dec)
    stack("decoder.", c["decoder_layer"], d_dec, c["decoder_head"])
    return d_enc, d_dec


BLOCK_SPECS = {"transformer": _transformer, "fastformer": _fastformer, "conformer": _conformer}

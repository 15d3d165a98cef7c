"""LSTM-with-dones over a short sequence (truncated BPTT) as a torch.autograd.Function whose pointwise work runs in the
hand-written kernels of csrc/vine_lstm.cu (one fused launch per step and direction); the GEMMs (input projection,
recurrent projection, their data/weight gradients) are plain bf16 library GEMMs.

Semantics == ppo.ActorCritic's reference loop (rl_games LSTMWithDones): the state entering step t is multiplied by
not_done[t]; gate order i, f, g, o; bias = b_ih + b_hh.
"""
import ctypes as C

import torch

from .. import abi


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class LstmSeqFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inp, w_ih, w_hh, bias, h0, c0, not_done):
        lib = abi.load_library()
        L, S, I = inp.shape
        H = w_hh.shape[1]
        stream = C.c_void_p(torch.cuda.current_stream(inp.device).cuda_stream)
        wih, whh = w_ih.detach().bfloat16(), w_hh.detach().bfloat16()
        x2 = inp.detach().reshape(L * S, I).bfloat16().contiguous()
        g_in = torch.addmm(bias.detach().bfloat16(), x2, wih.t()).view(L, S, 4 * H)
        nd = not_done.detach().float().contiguous() if not_done is not None else None
        dev = inp.device
        hs = torch.empty(L, S, H, device=dev)
        cs = torch.empty(L, S, H, device=dev)
        acts = torch.empty(L, S, 4 * H, device=dev, dtype=torch.bfloat16)
        h_ins = torch.empty(L, S, H, device=dev, dtype=torch.bfloat16)     # masked recurrent inputs of every step
        h_ins[0] = (h0.detach() * nd[0].unsqueeze(-1)) if nd is not None else h0.detach()
        c0 = c0.detach().float().contiguous()
        for t in range(L):
            gates = torch.addmm(g_in[t], h_ins[t], whh.t())               # bf16 [S, 4H]
            last = t == L - 1
            rc = lib.vine_lstm_cell_fwd(_p(gates), _p(cs[t - 1] if t else c0), _p(nd[t]) if nd is not None else None,
                                        _p(nd[t + 1]) if (nd is not None and not last) else None, S, H, _p(cs[t]), _p(hs[t]),
                                        None if last else _p(h_ins[t + 1]), _p(acts[t]), stream)
            assert rc == 0, rc
        ctx.save_for_backward(x2, wih, whh, acts, cs, c0, h_ins, nd if nd is not None else torch.empty(0, device=dev))
        ctx.has_nd = nd is not None
        ctx.dims = (L, S, I, H)
        ctx.in_dtype = inp.dtype
        return hs, hs[L - 1], cs[L - 1]

    @staticmethod
    def backward(ctx, dhs, dh_last, dc_last):
        lib = abi.load_library()
        x2, wih, whh, acts, cs, c0, h_ins, nd = ctx.saved_tensors
        nd = nd if ctx.has_nd else None
        L, S, I, H = ctx.dims
        dev = x2.device
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        dhs = dhs.float().contiguous() if dhs is not None else torch.zeros(L, S, H, device=dev)
        if dh_last is not None:
            dhs = dhs.clone()
            dhs[L - 1] += dh_last.float()
        dG = torch.empty(L, S, 4 * H, device=dev, dtype=torch.bfloat16)
        dc_bufs = [torch.empty(S, H, device=dev), torch.empty(S, H, device=dev)]
        dc_next = dc_last.float().contiguous() if dc_last is not None else None
        dh_rec = None
        for t in range(L - 1, -1, -1):
            out = dc_bufs[t & 1]
            rc = lib.vine_lstm_cell_bwd(_p(acts[t]), _p(cs[t - 1] if t else c0), _p(cs[t]), _p(nd[t]) if nd is not None else None,
                                        _p(dhs[t]), _p(dh_rec), _p(nd[t + 1]) if (nd is not None and t < L - 1) else None,
                                        _p(dc_next), S, H, _p(dG[t]), _p(out), stream)
            assert rc == 0, rc
            dh_rec = torch.mm(dG[t], whh)                                  # raw d(h_in_t); masked by the consumer with nd[t]
            dc_next = out
        dG2 = dG.view(L * S, 4 * H)
        d_whh = torch.mm(dG2.t(), h_ins.view(L * S, H)).float()
        d_wih = torch.mm(dG2.t(), x2).float()
        d_bias = dG2.float().sum(0)
        d_inp = torch.mm(dG2, wih).view(L, S, I).to(ctx.in_dtype)
        d_h0 = dh_rec.float() * nd[0].unsqueeze(-1) if nd is not None else dh_rec.float()
        return d_inp, d_wih, d_whh, d_bias, d_h0, dc_next, None


def lstm_seq(inp, w_ih, w_hh, bias, h0, c0, not_done=None):
    """inp [L, S, I]; returns (h of every step [L, S, H] f32, (h_L, c_L))."""
    hs, h_last, c_last = LstmSeqFunction.apply(inp, w_ih, w_hh, bias, h0, c0, not_done)
    return hs, (h_last, c_last)

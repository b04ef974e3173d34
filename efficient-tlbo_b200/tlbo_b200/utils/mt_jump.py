"""Jump-ahead polynomials for MT19937 (host side of the multi-CTA tie-break key generator).

MT19937's state transition is linear over GF(2): every bit of its word sequence x_0, x_1, ... satisfies the
linear recurrence whose characteristic polynomial phi(t) has degree 19937.  The window of 624 consecutive
words that starts J words later is therefore ``g_J(F) W`` with ``g_J(t) = t^J mod phi(t)`` and F the one-word
shift, i.e. word j of the jumped window is the XOR over the set bits i of g_J of x_{i + j} (Haramoto, Matsumoto,
Nishimura, Panneton, L'Ecuyer 2008, "Efficient jump ahead for F2-linear random number generators").  The
device evaluates that XOR (``tlbo_mt19937_jump``); this module supplies the bit positions:

* ``charpoly()``: phi by Berlekamp-Massey over 2 x 19937 + 64 output bits of a seeded numpy generator (any
  non-zero linear functional of the state has phi as its minimal polynomial; bit 0 of the tempered output is
  one), checked against the recurrence on a second stream;
* ``jump_poly(J)``: t^J mod phi by square-and-multiply on Python integers (bit i = coefficient of t^i).

Polynomials are Python ints; a carry-less product is a loop of shifted XORs (about 10 ms per product)."""
import numpy as np

DEG = 19937
_PHI = None


def _output_bits(seed, count):
    """Bit 0 of `count` consecutive 32-bit outputs of numpy's legacy MT19937."""
    raw = np.frombuffer(np.random.RandomState(seed).bytes(4 * count), dtype='<u4')
    return (raw & 1).astype(np.uint8)


def _berlekamp_massey(bits):
    """Connection polynomial C (int, bit i = c_i, c_0 = 1) and linear complexity L of a GF(2) sequence:
    s_n = sum_{i=1..L} c_i s_{n-i}."""
    C, B, L, m = 1, 1, 0, 1
    srev = 0                       # bit i = s_{n-i} once s_n has been pushed
    for n, s in enumerate(bits):
        srev = (srev << 1) | int(s)
        d = (C & srev).bit_count() & 1
        if d == 0:
            m += 1
        elif 2 * L <= n:
            T = C
            C ^= B << m
            L = n + 1 - L
            B = T
            m = 1
        else:
            C ^= B << m
            m += 1
        # keep only the bits a polynomial of degree <= L + m can reach
        if n % 4096 == 0:
            srev &= (1 << (DEG + 64)) - 1
    return C, L


def _reverse(C, L):
    """t^L C(1/t)."""
    out = 0
    for i in range(L + 1):
        if (C >> i) & 1:
            out |= 1 << (L - i)
    return out


def charpoly():
    """phi(t) of MT19937 as an int: sum_i phi_i x_{n+i} = 0 for every n and every bit of the word sequence."""
    global _PHI
    if _PHI is None:
        bits = _output_bits(4357, 2 * DEG + 64)
        C, L = _berlekamp_massey(bits)
        if L != DEG:
            raise RuntimeError('Berlekamp-Massey found linear complexity %d, expected %d' % (L, DEG))
        phi = _reverse(C, L)
        # check on a second stream: the recurrence must annihilate it
        chk = _output_bits(20261020, DEG + 2048)
        word = 0
        for i, b in enumerate(chk):
            word |= int(b) << i
        for n in (0, 1, 777, 2047):
            if ((word >> n) & phi).bit_count() & 1:
                raise RuntimeError('characteristic polynomial fails the recurrence check')
        _PHI = phi
    return _PHI


def _mulmod(a, b, phi):
    """Carry-less product a * b reduced mod phi."""
    acc = 0
    while a:
        low = a & -a
        acc ^= b << (low.bit_length() - 1)
        a ^= low
    return _mod(acc, phi)


def _mod(a, phi):
    dphi = phi.bit_length() - 1
    while True:
        da = a.bit_length() - 1
        if da < dphi:
            return a
        a ^= phi << (da - dphi)


def jump_poly(J):
    """t^J mod phi(t) as an int."""
    phi = charpoly()
    result, base = 1, 2          # 1, t
    e = int(J)
    while e:
        if e & 1:
            result = _mulmod(result, base, phi)
        e >>= 1
        if e:
            base = _mulmod(base, base, phi)
    return result


def bit_positions(poly):
    """Sorted positions of the set bits of a polynomial (int32 array): what the device kernel iterates over."""
    out = []
    i = 0
    while poly:
        low = poly & -poly
        out.append(low.bit_length() - 1)
        poly ^= low
    return np.array(out, dtype=np.int32)


# ---- numpy statement of the generator's word sequence (tests and the device kernel's specification) ---------
def twist_block(key):
    """Next 624-word block of the MT19937 word sequence (numpy's ``genrand`` regeneration, in place order)."""
    k = np.array(key, dtype=np.uint64)
    out = k.copy()
    for i in range(624):
        y = (out[i] & 0x80000000) | (out[(i + 1) % 624] & 0x7fffffff)
        out[i] = out[(i + 397) % 624] ^ (y >> np.uint64(1)) ^ (np.uint64(0x9908b0df) if (y & np.uint64(1)) else np.uint64(0))
    return out.astype(np.uint32)


def jump_window(key, positions):
    """``g(F) W`` for the window ``key`` (uint32[624]) and the set-bit positions of g: the window that starts J
    words later in the sequence (J the jump the polynomial was built for)."""
    blocks = [np.array(key, dtype=np.uint32)]
    need = int(positions.max()) + 624 if len(positions) else 624
    while 624 * len(blocks) < need:
        blocks.append(twist_block(blocks[-1]))
    seq = np.concatenate(blocks)
    out = np.zeros(624, dtype=np.uint32)
    for p in positions:
        out ^= seq[p:p + 624]
    return out

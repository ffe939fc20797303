"""Bit-string helpers with the reference's names and conventions (reference code/binary_io.py:7-67).

The grouped coders produce their bitstreams on the GPU (cwq_elias_delta_pack / cwq_fixed_pack); these
host versions serve the per-group entry points and callers that want a single code word.  They use
integer arithmetic (bit_length) where the reference uses float logs; the two agree for every value
the coders can produce (x <= 2**31)."""


def to_bit_string(num, num_bits):
    """Fixed-width LSB-first binary: to_bit_string(6, 4) == '0110' (BIO:41-53)."""
    num = int(num)
    num_bits = int(num_bits)
    if num >= 2 ** num_bits:
        raise Exception("The number {} (>= {}) is bigger than what we can encode!".format(num, 2 ** num_bits))
    return ''.join('1' if (num >> i) & 1 else '0' for i in range(num_bits))


def from_bit_string(bitcode):
    """Inverse of to_bit_string; accepts str or bytes (BIO:55-67)."""
    if isinstance(bitcode, bytes):
        bitcode = bitcode.decode("utf-8")
    num = 0
    for i, ch in enumerate(bitcode):
        if ch == "1":
            num += 1 << i
    return num


def elias_delta_code(x):
    """Elias-delta code of x >= 1 as a '0'/'1' string, every field MSB first (BIO:7-20)."""
    x = int(x)
    if x < 1:
        raise Exception("Elias delta codes positive integers only, got {}".format(x))
    n = x.bit_length() - 1            # floor(log2 x)
    l = (n + 1).bit_length() - 1      # floor(log2 (n + 1))
    length_code = format(n + 1, "b").zfill(l + 1)
    num_code = format(x, "b")[1:]
    return "0" * l + length_code + num_code


def elias_delta_decode(x):
    """First code word of a '0'/'1' byte string -> (value, code length) (BIO:22-39)."""
    if isinstance(x, str):
        x = x.encode("utf-8")
    l = 0
    while x[l] == 48:
        l += 1
    n_plus_one = int(x[l:2 * l + 1].decode("utf-8"), 2)
    rest = x[2 * l + 1:2 * l + n_plus_one].decode("utf-8")
    num = int("1" + rest, 2)
    return num, 2 * l + n_plus_one

"""Bit-string helpers with the reference's names and conventions (reference code/binary_io.py:7-67).

The grouped coders produce their bitstreams on the GPU (cwq_elias_delta_pack / cwq_fixed_pack); these
host versions serve the per-group entry points and callers that want a single code word.  They use
integer arithmetic (bit_length) where the reference uses float logs; the two agree for every value
the coders can produce (x <= 2**31).

`write_bin_code` / `read_bin_code` are the `.miracle` container (BIO:69-197): big-endian 4-byte header
integers, 16-bit-length-prefixed bit blobs (the arithmetic-coded group sizes), length/width-prefixed lists
of LSB-first fixed-width items (outlier indices and values) and the zero-padded payload, every bit string
packed MSB-first into bytes.  Files are byte-identical to the reference's (tests/golden/container_*)."""
import numpy as np


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


def _pack_bits(bits):
    """'0'/'1' characters (str or list of single characters) -> bytes, MSB first, zero padded (BIO:76-78)."""
    if not isinstance(bits, str):
        bits = ''.join(bits)
    if not bits:
        return b""
    a = np.frombuffer(bits.encode("ascii"), dtype=np.uint8) - np.uint8(48)
    return np.packbits(a).tobytes()


def _length_field(n, extra_bytes):
    if extra_bytes == 2:
        return bytes([n // 256, n % 256])       # BIO:97,110: raises above 65535 exactly like the reference
    return int(n).to_bytes(extra_bytes, "big")


def write_bin_code(code, path, extras=None, extra_var_bits=None, var_length_extras=None, var_length_bits=None, extra_bytes=2):
    """BIO:69-133.  extra_bytes: width of the length fields in front of the arithmetic-coded blobs and the outlier lists.
    The reference's writer hard-codes 2 (BIO:97,110) although its reader takes `extra_bytes` (BIO:136); with the default the
    file is byte-identical to the reference's.  A 3840x2160 image can hold more than 65 535 outliers / coded bits per blob,
    which the 16-bit fields cannot express: extra_bytes=4 is the documented extension, read back with
    read_bin_code(..., extra_bytes=4)."""
    if extra_bytes not in (2, 3, 4, 8):
        raise Exception("extra_bytes must be 2 (reference format), 3, 4 or 8")
    if var_length_extras is not None:
        if var_length_bits is None or len(var_length_extras) != len(var_length_bits):
            raise Exception("Each var length extra needs to have a bitlength associated!")
    with open(path, "wb") as f:
        if extras is not None:
            for extra in extras:
                f.write(int(extra % (256 ** 4)).to_bytes(4, "big"))      # BIO:84-92 keeps the low 4 bytes
        if extra_var_bits is not None:
            for bits in extra_var_bits:
                f.write(_length_field(len(bits), extra_bytes))
                f.write(_pack_bits(bits))
        if var_length_extras is not None:
            for extra, width in zip(var_length_extras, var_length_bits):
                f.write(_length_field(len(extra), extra_bytes))
                f.write(bytes([width]))                                 # BIO:113
                f.write(_pack_bits(''.join(to_bit_string(item, width) for item in extra)))
        f.write(_pack_bits(code))


def read_bin_code(path, num_extras=0, num_extra_var_bits=0, num_var_length_extras=0, extras_bytes=4, extra_bytes=2):
    """BIO:136-197 -> (payload bit string incl. padding, extras, extra_var_bits, var_length_extras)."""
    with open(path, "rb") as f:
        raw = np.frombuffer(f.read(), dtype=np.uint8)
    pos = 0
    extras = [int.from_bytes(raw[pos + i * extras_bytes:pos + (i + 1) * extras_bytes].tobytes(), "big") for i in range(num_extras)]
    pos += num_extras * extras_bytes

    def bit_chars(b0, n_bytes):
        return (np.unpackbits(raw[b0:b0 + n_bytes]) + np.uint8(48)).tobytes().decode("ascii")

    extra_var_bits = []
    for _ in range(num_extra_var_bits):
        n_bits = int.from_bytes(raw[pos:pos + extra_bytes].tobytes(), "big")
        pos += extra_bytes
        n_bytes = (n_bits + 7) // 8
        extra_var_bits.append(bit_chars(pos, n_bytes)[:n_bits])
        pos += n_bytes
    var_length_extras = []
    for _ in range(num_var_length_extras):
        length = int.from_bytes(raw[pos:pos + extra_bytes].tobytes(), "big")
        width = int(raw[pos + extra_bytes])
        pos += extra_bytes + 1
        n_bytes = (width * length + 7) // 8
        chars = bit_chars(pos, n_bytes)
        var_length_extras.append([from_bit_string(chars[i * width:(i + 1) * width]) for i in range(length)])
        pos += n_bytes
    return bit_chars(pos, raw.size - pos), extras, extra_var_bits, var_length_extras

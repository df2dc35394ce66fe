"""Section parser for the .nhw container (writer: reference
encoder/nhw_encoder.c:3100-3220, reader: decoder/nhw_decoder.c:1494-1661).  Host-side tooling:
used by tests and debugging to name the first section in which two streams differ."""
import struct


def parse_nhw(data):
    data = bytes(data)
    pos = 0

    def u8():
        nonlocal pos
        v = data[pos]
        pos += 1
        return v

    def u16():
        nonlocal pos
        v = struct.unpack_from("<H", data, pos)[0]
        pos += 2
        return v

    def u32():
        nonlocal pos
        v = struct.unpack_from("<I", data, pos)[0]
        pos += 4
        return v

    h = {}
    h["res_high_wavelet"] = u8()
    q = h["quality"] = u8()
    h["size_tree1"], h["size_tree2"] = u16(), u16()
    h["size_data1"], h["size_data2"] = u32(), u32()
    h["tree_end"], h["exw_Y_end"] = u16(), u16()
    if q > 12:
        h["res1_len"] = u16()
    if q >= 19:
        h["res3_len"], h["res3_bit_len"] = u16(), u16()
    if q > 17:
        h["res4_len"] = u16()
    if q > 12:
        h["res1_bit_len"] = u16()
    if q >= 21:
        h["res5_len"], h["res5_bit_len"] = u16(), u16()
    if q > 21:
        h["res6_len"], h["res6_bit_len"], h["char_res1_len"] = u32(), u16(), u16()
    if q > 22:
        h["qsetting3_len"] = u16()
    h["select1"], h["select2"] = u16(), u16()
    if q > 15:
        h["highres_comp_len"] = u16()
    h["end_ch_res"] = u16()
    h["header_bytes"] = pos
    sec = {}

    def take(name, n):
        nonlocal pos
        sec[name] = data[pos:pos + n]
        pos += n

    take("tree1", h["size_tree1"])
    take("tree2", h["size_tree2"])
    take("exw_Y", h["exw_Y_end"])
    if q > 12:
        take("res1", h["res1_len"])
        take("res1_bit", h["res1_bit_len"])
        take("res1_word", h["res1_bit_len"])
    if q > 17:
        take("res4", h["res4_len"])
    if q >= 19:
        take("res3", h["res3_len"])
        take("res3_bit", h["res3_bit_len"])
        take("res3_word", 2 * h["res3_bit_len"])
    if q >= 21:
        take("res5", h["res5_len"])
        take("res5_bit", h["res5_bit_len"])
        take("res5_word", h["res5_bit_len"])
    if q > 21:
        take("res6", h["res6_len"])
        take("res6_bit", h["res6_bit_len"])
        take("res6_word", h["res6_bit_len"])
        take("char_res1", 2 * h["char_res1_len"])
    if q > 22:
        take("high_qsetting3", 4 * h["qsetting3_len"])
    take("select_word1", h["select1"])
    take("select_word2", h["select2"])
    if q > 15:
        take("res_U_64", 512)
        take("res_V_64", 512)
        take("highres_word", h["highres_comp_len"])
    take("ch_res", h["end_ch_res"])
    take("encode", 4 * h["size_data2"])
    h["parsed_bytes"] = pos
    h["file_bytes"] = len(data)
    return h, sec


def first_difference(a, b):
    """-> human-readable description of the first header field / section in which a and b differ"""
    ha, sa = parse_nhw(a)
    hb, sb = parse_nhw(b)
    for k in ha:
        if ha.get(k) != hb.get(k):
            return "header field %s: %r vs %r" % (k, ha.get(k), hb.get(k))
    for k in sa:
        if sa[k] != sb.get(k):
            n = min(len(sa[k]), len(sb[k]))
            d = next((i for i in range(n) if sa[k][i] != sb[k][i]), n)
            return "section %s: first differing byte %d of %d" % (k, d, len(sa[k]))
    return None

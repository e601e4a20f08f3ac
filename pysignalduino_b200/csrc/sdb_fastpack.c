/*
 * sdb_fastpack.c — CPython extension: parser dicts -> packed SdbPulseMsg records + nibble-packed digit pool.
 *
 * The native counterpart of pack.pack_pulse (pysignalduino_b200/pack.py), i.e. of the input handling at the top of the
 * reference demodulators (sd_protocols/message_synced.py:21-66, message_unsynced.py:22-35), for the dict-level batch API
 * (SDProtocols.demodulate_batch): the Python loop packs 60-80 k dicts/s, three orders of magnitude below the kernels
 * behind it.  This module walks the dicts with the C API.
 *
 * Exactness rule: the fast path only handles what it can reproduce exactly — exact `dict` messages whose keys / values are
 * exact `str`, ASCII, pattern values in the canonical integer syntax -?[0-9]{1,10} (or empty = skipped slot).  Anything else
 * (values float() reads differently such as "+330", " 330", "3e2", "1_0"; non-str values; non-ASCII text) makes
 * pack_pulse() return None and the caller falls back to the Python packer for the whole batch — never a different result.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

#define MAX_SLOTS 8
#define MAX_DIGITS 4096
#define MSG_VALID 0x01
#define MSG_DOMAIN 0x04

typedef struct {
    int32_t pat[8];
    uint32_t doff;
    uint16_t dlen;
    uint8_t npat, cp;
    uint32_t pat_ids;
    uint8_t flags, rsv[3];
} PulseMsg;

static PyObject *s_data, *s_CP, *s_SP, *s_R;

/* ASCII exact-str view; returns 0 when the object is not an exact ASCII str */
static int ascii_view(PyObject *o, const char **p, Py_ssize_t *n)
{
    if (!PyUnicode_CheckExact(o) || !PyUnicode_IS_ASCII(o)) return 0;
    *p = (const char *)PyUnicode_1BYTE_DATA(o);
    *n = PyUnicode_GET_LENGTH(o);
    return 1;
}
static int all_digits(const char *p, Py_ssize_t n)
{
    if (n <= 0) return 0;
    for (Py_ssize_t i = 0; i < n; i++) if (p[i] < '0' || p[i] > '9') return 0;
    return 1;
}

/* pack_pulse(msgs: list, kind: int, rec: writable buffer of n * 48 bytes, clock: writable buffer of n doubles)
 *   -> None (fall back to the Python packer) | (pool: bytes, rssi: list, domain: dict) */
static PyObject *fast_pack_pulse(PyObject *self, PyObject *args)
{
    PyObject *msgs;
    int kind;
    Py_buffer recb, clkb;
    if (!PyArg_ParseTuple(args, "O!iw*w*", &PyList_Type, &msgs, &kind, &recb, &clkb)) return NULL;
    const Py_ssize_t n = PyList_GET_SIZE(msgs);
    PyObject *result = NULL, *rssi = NULL, *domain = NULL, *pool = NULL;
    int fallback = 0;
    if (recb.len < n * (Py_ssize_t)sizeof(PulseMsg) || clkb.len < n * (Py_ssize_t)sizeof(double)) {
        PyErr_SetString(PyExc_ValueError, "fast_pack_pulse: output buffers too small");
        goto done;
    }
    PulseMsg *rec = (PulseMsg *)recb.buf;
    double *clock = (double *)clkb.buf;
    const int is_ms = kind == 0;
    memset(rec, 0, (size_t)n * sizeof(PulseMsg));
    rssi = PyList_New(n);
    domain = PyDict_New();
    if (!rssi || !domain) goto done;

    /* pass 1: gates, patterns, record fields; digit units are summed for the pool */
    uint64_t units = 0;
    for (Py_ssize_t i = 0; i < n && !fallback; i++) {
        PyObject *m = PyList_GET_ITEM(msgs, i);
        PulseMsg *r = &rec[i];
        r->cp = 0xFF;
        r->doff = (uint32_t)units;                                    /* running offset, also for messages without digits */
        clock[i] = 0.0;
        if (!PyDict_CheckExact(m)) { fallback = 1; break; }
        PyObject *rv = PyDict_GetItemWithError(m, s_R);               /* rssi.append(m.get("R")) */
        if (!rv && PyErr_Occurred()) goto done;
        if (!rv) rv = Py_None;
        Py_INCREF(rv);
        PyList_SET_ITEM(rssi, i, rv);
        PyObject *dv = PyDict_GetItemWithError(m, s_data);
        if (!dv && PyErr_Occurred()) goto done;
        const char *dp = ""; Py_ssize_t dn = 0;
        if (dv && !ascii_view(dv, &dp, &dn)) { fallback = 1; break; }
        int valid;
        if (is_ms) {                                                  /* message_synced.py:21-47 */
            valid = all_digits(dp, dn);
            const char *q; Py_ssize_t qn;
            PyObject *cpv = PyDict_GetItemWithError(m, s_CP), *spv = PyDict_GetItemWithError(m, s_SP);
            if (PyErr_Occurred()) goto done;
            if (cpv && !ascii_view(cpv, &q, &qn)) { fallback = 1; break; }
            if (!cpv || !all_digits(q, qn)) valid = 0;
            if (spv && !ascii_view(spv, &q, &qn)) { fallback = 1; break; }
            if (!spv || !all_digits(q, qn)) valid = 0;
            if (PyDict_Contains(m, s_R) == 1) {                       /* "R" in msg_data: must be a digit string */
                if (!ascii_view(rv, &q, &qn)) { fallback = 1; break; }
                if (!all_digits(q, qn)) valid = 0;
            }
        } else valid = dn > 0;                                        /* message_unsynced.py:22-25 */
        if (!valid) continue;
        /* parse_patterns: P<digits> keys in dict order, id = int(key[1:]), later duplicates overwrite the value in place */
        int64_t ids[64]; int64_t vals[64]; int np = 0;
        Py_ssize_t pos = 0; PyObject *k, *v;
        while (PyDict_Next(m, &pos, &k, &v)) {
            const char *kp; Py_ssize_t kn;
            if (!PyUnicode_CheckExact(k)) { fallback = 1; break; }
            if (!PyUnicode_IS_ASCII(k)) {
                /* a non-ASCII key is a pattern key only if it starts with 'P' and the rest isdigit() in the Unicode sense */
                if (PyUnicode_GET_LENGTH(k) > 0 && PyUnicode_READ_CHAR(k, 0) == 'P') { fallback = 1; break; }
                continue;
            }
            kp = (const char *)PyUnicode_1BYTE_DATA(k); kn = PyUnicode_GET_LENGTH(k);
            if (kn < 2 || kp[0] != 'P' || !all_digits(kp + 1, kn - 1)) continue;
            const char *vp; Py_ssize_t vn;
            if (!ascii_view(v, &vp, &vn)) { fallback = 1; break; }
            if (vn == 0) continue;                                    /* float('') raises ValueError: slot skipped */
            /* id: leading zeros dropped; more than 18 digits cannot be compared cheaply -> fallback */
            const char *ip = kp + 1; Py_ssize_t in_ = kn - 1;
            while (in_ > 1 && *ip == '0') { ip++; in_--; }
            if (in_ > 18) { fallback = 1; break; }
            int64_t id = 0;
            for (Py_ssize_t j = 0; j < in_; j++) id = id * 10 + (ip[j] - '0');
            /* value: canonical integer only */
            Py_ssize_t a = 0; int neg = 0;
            if (vp[0] == '-') { neg = 1; a = 1; }
            if (vn - a < 1 || vn - a > 10 || !all_digits(vp + a, vn - a)) { fallback = 1; break; }
            int64_t val = 0;
            for (Py_ssize_t j = a; j < vn; j++) val = val * 10 + (vp[j] - '0');
            if (neg) val = -val;
            int slot = -1;
            for (int s = 0; s < np; s++) if (ids[s] == id) { slot = s; break; }
            if (slot < 0) { if (np == 64) { fallback = 1; break; } slot = np++; ids[slot] = id; }
            vals[slot] = val;
        }
        if (fallback) break;
        const char *why = NULL;
        char whybuf[96];
        if (dn > MAX_DIGITS) { snprintf(whybuf, sizeof whybuf, "D has %zd digits (max %d)", dn, MAX_DIGITS); why = whybuf; }
        else if (np > MAX_SLOTS) { snprintf(whybuf, sizeof whybuf, "%d pattern slots (max %d)", np, MAX_SLOTS); why = whybuf; }
        else for (int s = 0; s < np && !why; s++) {
            if (ids[s] > 9) { snprintf(whybuf, sizeof whybuf, "pattern id '%lld' is not a single digit", (long long)ids[s]); why = whybuf; }
            else if (vals[s] > 2147483647LL || vals[s] < -2147483647LL) { snprintf(whybuf, sizeof whybuf, "pattern value %lld.0 is not an int32", (long long)vals[s]); why = whybuf; }
        }
        if (why) {
            PyObject *key = PyLong_FromSsize_t(i), *txt = PyUnicode_FromString(why);
            if (!key || !txt || PyDict_SetItem(domain, key, txt) < 0) { Py_XDECREF(key); Py_XDECREF(txt); goto done; }
            Py_DECREF(key); Py_DECREF(txt);
            r->flags = MSG_DOMAIN;
            continue;
        }
        uint32_t pid = 0;
        for (int s = 0; s < np; s++) { r->pat[s] = (int32_t)vals[s]; pid |= (uint32_t)ids[s] << (4 * s); }
        r->pat_ids = pid; r->npat = (uint8_t)np; r->dlen = (uint16_t)dn; r->flags = MSG_VALID;
        if (is_ms) {                                                  /* cp_key = str(int(CP)) looked up among the pattern ids */
            const char *q; Py_ssize_t qn;
            ascii_view(PyDict_GetItem(m, s_CP), &q, &qn);
            while (qn > 1 && *q == '0') { q++; qn--; }
            if (qn == 1) {
                const int cpid = q[0] - '0';
                for (int s = 0; s < np; s++) if (ids[s] == cpid) { r->cp = (uint8_t)s; clock[i] = (double)(vals[s] < 0 ? -vals[s] : vals[s]); break; }
            }
        }
        units += (uint64_t)((dn + 31) / 32);
    }
    if (fallback) { result = Py_None; Py_INCREF(result); goto done; }
    if (units * 16 + 64 > 0xFFFFFFFFull * 16ull) { PyErr_SetString(PyExc_OverflowError, "digit pool too large"); goto done; }

    /* pass 2: the digit pool (one nibble per character, 0xE for non-digits, 0xF padding to the 16-byte unit + 32 B tail) */
    pool = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)(units * 16 + 32));
    if (!pool) goto done;
    uint8_t *pp = (uint8_t *)PyBytes_AS_STRING(pool);
    memset(pp, 0xFF, (size_t)(units * 16 + 32));
    for (Py_ssize_t i = 0; i < n; i++) {
        const PulseMsg *r = &rec[i];
        if (!(r->flags & MSG_VALID)) continue;
        PyObject *dv = PyDict_GetItem(PyList_GET_ITEM(msgs, i), s_data);
        const uint8_t *dp = (const uint8_t *)PyUnicode_1BYTE_DATA(dv);
        const Py_ssize_t dn = r->dlen;
        uint8_t *dst = pp + (size_t)r->doff * 16;
        Py_ssize_t j = 0;
        for (; j + 1 < dn; j += 2) {
            uint8_t a = (uint8_t)(dp[j] - '0'), b = (uint8_t)(dp[j + 1] - '0');
            if (a > 9) a = 0xE;
            if (b > 9) b = 0xE;
            dst[j >> 1] = (uint8_t)(a | (b << 4));
        }
        if (j < dn) { uint8_t a = (uint8_t)(dp[j] - '0'); if (a > 9) a = 0xE; dst[j >> 1] = (uint8_t)(a | 0xF0); }
    }
    result = PyTuple_Pack(3, pool, rssi, domain);
done:
    Py_XDECREF(pool); Py_XDECREF(rssi); Py_XDECREF(domain);
    PyBuffer_Release(&recb); PyBuffer_Release(&clkb);
    return result;
}

static PyMethodDef methods[] = {
    {"pack_pulse", fast_pack_pulse, METH_VARARGS, "pack a list of MS / MU parser dicts (None = use the Python packer)"},
    {NULL, NULL, 0, NULL},
};
static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_fastpack", "native dict packer of pysignalduino_b200", -1, methods};

PyMODINIT_FUNC PyInit__fastpack(void)
{
    s_data = PyUnicode_InternFromString("data"); s_CP = PyUnicode_InternFromString("CP");
    s_SP = PyUnicode_InternFromString("SP"); s_R = PyUnicode_InternFromString("R");
    return PyModule_Create(&moddef);
}

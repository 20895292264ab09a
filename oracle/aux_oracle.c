/*
 * aux_oracle.c -- CPU restatements of the two callers' data formats either side of the scan.
 *
 * TEST INFRASTRUCTURE ONLY (see cortex_oracle.c): the checker for cx_extract_embeddings /
 * cx_load_nodes and cx_apply_score_decay.
 *
 *   cxo_walk_node          bincode 1.3 (fixint, little endian, trailing bytes allowed) of `Node`
 *                          /root/reference/crates/cortex-core/src/types.rs:26-68 (Node), :126-145
 *                          (NodeData), :274-283 (Source); layout PINNED by the reference's golden
 *                          bytes, storage/redb_storage.rs:1834-1856 (tests/golden/node_golden.bin)
 *   cxo_apply_score_decay  /root/reference/crates/cortex-core/src/vector/scoring.rs:84-114,
 *                          pinned by the assertions of its unit tests (:136-260)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

enum { NODE_OK = 0, NODE_NO_EMBEDDING = 1, NODE_DELETED = 2, NODE_DIM_MISMATCH = 3, NODE_NEEDS_HOST_DECODE = 4, NODE_CORRUPT = 5 };

typedef struct {
  const uint8_t *p;
  size_t off, end;
  int bad;
} cur;

static uint8_t rd8(cur *c) {
  if (c->off + 1 > c->end) {
    c->bad = 1;
    return 0;
  }
  return c->p[c->off++];
}
static uint64_t rd64(cur *c) {
  if (c->off + 8 > c->end) {
    c->bad = 1;
    return 0;
  }
  uint64_t v;
  memcpy(&v, c->p + c->off, 8); /* little-endian host */
  c->off += 8;
  return v;
}

/* String::from_utf8 semantics */
static int utf8_ok(const uint8_t *s, size_t n) {
  size_t i = 0;
  while (i < n) {
    uint8_t c = s[i];
    if (c < 0x80) {
      ++i;
      continue;
    }
    uint32_t need, min_cp, cp;
    if ((c & 0xE0) == 0xC0) need = 1, min_cp = 0x80, cp = c & 0x1F;
    else if ((c & 0xF0) == 0xE0) need = 2, min_cp = 0x800, cp = c & 0x0F;
    else if ((c & 0xF8) == 0xF0) need = 3, min_cp = 0x10000, cp = c & 0x07;
    else return 0;
    if (i + need >= n) return 0;
    for (uint32_t k = 1; k <= need; ++k) {
      uint8_t x = s[i + k];
      if ((x & 0xC0) != 0x80) return 0;
      cp = (cp << 6) | (x & 0x3F);
    }
    if (cp < min_cp || cp > 0x10FFFF || (cp >= 0xD800 && cp <= 0xDFFF)) return 0;
    i += need + 1;
  }
  return 1;
}

static int rd_string(cur *c, const uint8_t **s, size_t *n) {
  uint64_t len = rd64(c);
  if (c->bad || len > c->end - c->off) {
    c->bad = 1;
    return 0;
  }
  if (!utf8_ok(c->p + c->off, (size_t)len)) {
    c->bad = 1;
    return 0;
  }
  if (s) *s = c->p + c->off;
  if (n) *n = (size_t)len;
  c->off += (size_t)len;
  return 1;
}
static int rd_opt_string(cur *c) {
  uint8_t tag = rd8(c);
  if (c->bad || tag > 1) {
    c->bad = 1;
    return 0;
  }
  return tag ? rd_string(c, NULL, NULL) : 1;
}

static int64_t days_from_civil(int64_t y, int64_t m, int64_t d) {
  y -= m <= 2;
  int64_t era = (y >= 0 ? y : y - 399) / 400;
  int64_t yoe = y - era * 400;
  int64_t doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
  int64_t doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
  return era * 146097 + doe - 719468;
}

static int num(const uint8_t *s, size_t n, size_t i, int k, int64_t *v) {
  int64_t x = 0;
  for (int j = 0; j < k; ++j) {
    if (i + j >= n || s[i + j] < '0' || s[i + j] > '9') return 0;
    x = x * 10 + (s[i + j] - '0');
  }
  *v = x;
  return 1;
}

/* the RFC 3339 subset chrono's serializer emits (plus numeric offsets); anything else -> 0 */
static int parse_rfc3339(const uint8_t *s, size_t n, int64_t *out_ns) {
  int64_t Y, M, D, h, m, sec, frac = 0, off_s = 0;
  if (n < 20) return 0;
  if (!num(s, n, 0, 4, &Y) || s[4] != '-' || !num(s, n, 5, 2, &M) || s[7] != '-' || !num(s, n, 8, 2, &D)) return 0;
  if (s[10] != 'T' || !num(s, n, 11, 2, &h) || s[13] != ':' || !num(s, n, 14, 2, &m) || s[16] != ':' ||
      !num(s, n, 17, 2, &sec))
    return 0;
  if (M < 1 || M > 12 || D < 1 || D > 31 || h > 23 || m > 59 || sec > 59) return 0;
  size_t i = 19;
  if (i < n && s[i] == '.') {
    int nd = 0;
    ++i;
    while (i < n && s[i] >= '0' && s[i] <= '9') {
      if (nd < 9) frac = frac * 10 + (s[i] - '0');
      ++nd;
      ++i;
    }
    if (nd == 0 || nd > 9) return 0;
    for (; nd < 9; ++nd) frac *= 10;
  }
  if (i < n && s[i] == 'Z') {
    ++i;
  } else if (i < n && (s[i] == '+' || s[i] == '-')) {
    int64_t sign = s[i] == '-' ? -1 : 1, oh, om;
    if (!num(s, n, i + 1, 2, &oh) || i + 3 >= n || s[i + 3] != ':' || !num(s, n, i + 4, 2, &om) || oh > 23 || om > 59)
      return 0;
    off_s = sign * (oh * 3600 + om * 60);
    i += 6;
  } else {
    return 0;
  }
  if (i != n) return 0;
  int64_t secs = days_from_civil(Y, M, D) * 86400 + h * 3600 + m * 60 + sec - off_s;
  if (secs > 9000000000ll || secs < -9000000000ll) return 0;
  *out_ns = secs * 1000000000ll + frac;
  return 1;
}

/* Walk one serialized Node in field order (types.rs:26-68).  Returns the status; id[16], the embedding
 * (copied to out_row if status == NODE_OK and out_row != NULL) and the timestamps / access count. */
int cxo_walk_node(const uint8_t *v, size_t len, size_t dim, uint8_t *id, float *out_row, int64_t *created_ns,
                  int64_t *last_accessed_ns, uint64_t *access_count) {
  cur c = {v, 0, len, 0};
  if (rd64(&c) != 16 || c.bad) return NODE_CORRUPT; /* Uuid: byte sequence of length 16 */
  if (c.off + 16 > c.end) return NODE_CORRUPT;
  memcpy(id, c.p + c.off, 16);
  c.off += 16;
  if (!rd_string(&c, NULL, NULL)) return NODE_CORRUPT; /* kind: NodeKind(String) */
  if (!rd_string(&c, NULL, NULL)) return NODE_CORRUPT; /* data.title */
  if (!rd_string(&c, NULL, NULL)) return NODE_CORRUPT; /* data.body */
  uint64_t n_meta = rd64(&c);                          /* data.metadata: HashMap<String, serde_json::Value> */
  if (c.bad) return NODE_CORRUPT;
  if (n_meta != 0) return NODE_NEEDS_HOST_DECODE;
  uint64_t n_tags = rd64(&c); /* data.tags: Vec<String> */
  if (c.bad || n_tags > (c.end - c.off) / 8) return NODE_CORRUPT;
  for (uint64_t t = 0; t < n_tags; ++t)
    if (!rd_string(&c, NULL, NULL)) return NODE_CORRUPT;
  uint8_t has_emb = rd8(&c); /* embedding: Option<Vec<f32>> */
  if (c.bad || has_emb > 1) return NODE_CORRUPT;
  uint64_t emb_len = 0;
  size_t emb_off = 0;
  if (has_emb) {
    emb_len = rd64(&c);
    if (c.bad || emb_len > (c.end - c.off) / 4) return NODE_CORRUPT;
    emb_off = c.off;
    c.off += (size_t)emb_len * 4;
  }
  if (!rd_string(&c, NULL, NULL)) return NODE_CORRUPT; /* source.agent */
  if (!rd_opt_string(&c)) return NODE_CORRUPT;         /* source.session */
  if (!rd_opt_string(&c)) return NODE_CORRUPT;         /* source.channel */
  if (c.off + 4 > c.end) return NODE_CORRUPT;          /* importance: f32 */
  c.off += 4;
  *access_count = rd64(&c);
  if (c.bad) return NODE_CORRUPT;
  const uint8_t *ts[3];
  size_t tn[3];
  for (int t = 0; t < 3; ++t) /* last_accessed_at, created_at, updated_at: DateTime<Utc> as RFC 3339 strings */
    if (!rd_string(&c, &ts[t], &tn[t])) return NODE_CORRUPT;
  uint8_t deleted = rd8(&c);
  if (c.bad || deleted > 1) return NODE_CORRUPT;
  int64_t upd;
  if (!parse_rfc3339(ts[0], tn[0], last_accessed_ns) || !parse_rfc3339(ts[1], tn[1], created_ns) ||
      !parse_rfc3339(ts[2], tn[2], &upd))
    return NODE_NEEDS_HOST_DECODE;
  if (deleted) return NODE_DELETED;
  if (!has_emb) return NODE_NO_EMBEDDING;
  if (emb_len != dim) return NODE_DIM_MISMATCH;
  if (out_row) memcpy(out_row, v + emb_off, dim * 4);
  return NODE_OK;
}

/* vector/scoring.rs:84-114 */
float cxo_apply_score_decay(float raw_score, int64_t idle_seconds, uint64_t access_count, double kind_rate, int enabled,
                            double max_age_days, double min_factor, double echo_weight, double echo_cap,
                            float recency_bias) {
  if (!enabled || recency_bias == 0.0f) return raw_score;
  double days_idle = (double)(idle_seconds > 0 ? idle_seconds : 0) / 86400.0;
  double effective_days = fmin(days_idle, max_age_days);
  float temporal_factor = (float)fmax(exp(-kind_rate * effective_days), min_factor);
  float echo_factor = (float)fmin(1.0 + (double)access_count * echo_weight, echo_cap);
  /* raw * (1 - bias) + raw * temporal * echo * bias, left to right in f32 */
  volatile float keep = raw_score * (1.0f - recency_bias);
  volatile float m1 = raw_score * temporal_factor;
  volatile float m2 = m1 * echo_factor;
  volatile float m3 = m2 * recency_bias;
  return keep + m3;
}

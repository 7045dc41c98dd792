"""Host-side word / embedding containers of the reference's Python module: `WordDict` (py/src/variants/mod.rs:8-78) and
`Embeddings` (py/src/embeddings.rs:8-144).  Pure bookkeeping (dictionary lookups, ordered f32 row sums, file formats);
every distance goes through the GPU (granne_b200_compute_distances)."""
import json

import numpy as np


class WordDict:
    """py/src/variants/mod.rs:8-78: one JSON string per line; ids are line numbers."""

    def __init__(self, path=None):
        self.id_to_word = []
        self.word_to_id = {}
        if path is not None:
            with open(path, "r", encoding="utf-8") as f:
                for line in f.read().splitlines():
                    word = json.loads(line)
                    if not isinstance(word, str):
                        raise ValueError("words file: every line must be a JSON string")
                    self.id_to_word.append(word)
            # collect() into a HashMap keeps the LAST id of a duplicated word (:26)
            self.word_to_id = {w: i for i, w in enumerate(self.id_to_word)}

    def __len__(self):
        return len(self.id_to_word)

    def get_words(self, ids):
        return " ".join(self.id_to_word[i] for i in ids)  # :38-51

    def get_word_ids(self, query):
        return [self.word_to_id[w] for w in query.split() if w in self.word_to_id]  # :53-58 (unknown words dropped)

    def push(self, word):
        if word in self.word_to_id:  # :60-69
            return False
        self.word_to_id[word] = len(self.id_to_word)
        self.id_to_word.append(word)
        return True

    def write(self, path):
        with open(path, "w", encoding="utf-8") as f:
            for word in self.id_to_word:
                f.write(json.dumps(word, ensure_ascii=False) + "\n")  # serde_json::to_string keeps non-ASCII as is


def read_dense_f32(data):
    """FixedWidthSliceVector<f32> file image (src/slice_vector/mod.rs:213-221): u64 width + rows."""
    buf = np.frombuffer(data, dtype=np.uint8)
    if buf.size < 8:
        raise ValueError("embeddings file shorter than its width prefix")
    dim = int(np.frombuffer(buf[:8].tobytes(), dtype="<u8")[0])
    body = np.frombuffer(buf[8:].tobytes(), dtype="<f4")
    if dim == 0 or body.size % dim != 0:
        raise ValueError("embeddings file: width must be > 0 and divide the payload")
    return body.reshape(-1, dim).copy()


def read_sum_terms(data):
    """VariableWidthSliceVector<ThreeByteInt, FiveByteInt> image (src/slice_vector/mod.rs:660-676) -> list of id lists."""
    buf = np.frombuffer(data, dtype=np.uint8)
    n = int(np.frombuffer(buf[:8].tobytes(), dtype="<u8")[0]) if buf.size >= 8 else -1
    if n < 0 or 8 + (n + 1) * 5 > buf.size:
        raise ValueError("embeddings elements file: offset table exceeds the file")
    off = buf[8:8 + (n + 1) * 5].reshape(n + 1, 5).astype(np.uint64)
    offsets = sum(off[:, b] << np.uint64(8 * b) for b in range(5))
    body = buf[8 + (n + 1) * 5:]
    ids = body[:body.size // 3 * 3].reshape(-1, 3).astype(np.uint32)
    terms = ids[:, 0] | (ids[:, 1] << 8) | (ids[:, 2] << 16)
    return [terms[int(offsets[i]):int(offsets[i + 1])].tolist() for i in range(n)]


def write_sum_terms(lists):
    """The inverse of read_sum_terms: VariableWidthSliceVector<ThreeByteInt, FiveByteInt>::write
    (src/slice_vector/mod.rs:623-634): u64 count | (count + 1) 5-byte offsets | 3-byte ids."""
    out = bytearray(len(lists).to_bytes(8, "little"))
    total = 0
    out += total.to_bytes(5, "little")
    for l in lists:
        total += len(l)
        out += total.to_bytes(5, "little")
    for l in lists:
        for t in l:
            out += int(t).to_bytes(3, "little")
    return bytes(out)


def create_embedding(table, ids, dim=None):
    """SumEmbeddings::create_embedding (src/elements/embeddings/mod.rs:119-143): the first row, then ordered
    element-wise f32 adds of the others (sum_into_f32, src/math.rs:92-116); no ids -> zeros (or [] for an empty table)."""
    if len(ids) == 0:
        return np.zeros(table.shape[1] if table.shape[0] > 0 else 0, dtype=np.float32)
    data = table[ids[0]].astype(np.float32, copy=True)
    for w in ids[1:]:
        data += table[w]
    return data


class Embeddings:
    """granne.Embeddings (py/src/embeddings.rs:8-144)."""

    def __init__(self, embeddings_path=None, words_path=None, device=0):
        if (embeddings_path is None) != (words_path is None):
            raise ValueError("embeddings_path and words_path must be given together")  # :36-37
        self.device = device
        if embeddings_path is None:
            self._rows = []
            self._dim = None
            self.words = WordDict()
        else:
            with open(embeddings_path, "rb") as f:
                table = read_dense_f32(f.read())
            self._rows = [r for r in table]
            self._dim = table.shape[1]
            self.words = WordDict(words_path)

    def _table(self):
        if not self._rows:
            return np.zeros((0, self._dim or 0), dtype=np.float32)
        return np.stack(self._rows).astype(np.float32, copy=False)

    def __len__(self):
        return len(self._rows)

    def _ids(self, value):
        if isinstance(value, str):
            return self.words.get_word_ids(value)
        if isinstance(value, (int, np.integer)):
            return [int(value)]
        return [int(v) for v in value]

    def get_embedding(self, value):
        """The (non-normalised) embedding for a word / sentence or an id / list of ids (:63-75)."""
        return create_embedding(self._table(), self._ids(value)).tolist()

    def dists(self, left, rights):
        """Angular distance between `left` and every entry of `rights` (:86-95)."""
        from . import api

        if len(rights) == 0:
            return []
        l = np.asarray(self.get_embedding(left), dtype=np.float32)
        r = np.stack([np.asarray(self.get_embedding(x), dtype=np.float32) for x in rights])
        return api.compute_distances("angular", np.repeat(l[None, :], r.shape[0], axis=0), r, self.device).tolist()

    def dist(self, left, right):
        return self.dists(left, [right])[0]  # :78-83

    def append(self, embedding, word):
        """Appends an embedding with its word; False (and nothing stored) if the word exists (:109-116)."""
        inserted = self.words.push(word)
        if inserted:
            row = np.asarray(embedding, dtype=np.float32)
            if self._dim is None:
                self._dim = row.size
            if row.size != self._dim:
                raise ValueError("embedding width differs from the table's")  # push asserts data.len() == width
            self._rows.append(row)
        return inserted

    def save_embeddings(self, path):
        with open(path, "wb") as f:
            f.write(np.uint64(self._dim or 0).astype("<u8").tobytes())  # FixedWidthSliceVector::write, mod.rs:460-466
            f.write(self._table().astype("<f4").tobytes())

    def save_words(self, path):
        self.words.write(path)

    def save(self, embeddings_path, words_path):
        self.save_embeddings(embeddings_path)
        self.save_words(words_path)

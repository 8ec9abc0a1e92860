/**
 * Wiring of the reference's FHEEngineImpl (src/api/fhe-engine.ts:209-321) to the B200 addon.
 *
 * The reference's engine returns placeholder ciphertexts `{handle: 0n}` and does no arithmetic (SURVEY H1): its
 * `add / batchAdd / multiply / relinearize / bootstrap` only compute the metadata (keyId check, noiseBudget, degree).
 * This module keeps that metadata bookkeeping byte for byte and attaches the real words: every Ciphertext handle indexes
 * a store of BigUint64Array words ([2][N], or [3][N] after multiply), and the data-parallel methods call the addon
 * (addon/fheb_addon.d.ts), which calls libfheb200.so.  Key generation, encryption and decryption stay where the
 * reference has them (host code, out of this backend's scope); a caller that already holds ciphertext words - e.g.
 * ballots received as FHEV records - registers them with `importCiphertext`.
 *
 * Usage (drop-in): `const engine = wireB200(await createFHEEngine('tfhe-128-fast'), require('./build/Release/node-fhe-accelerate.node'))`.
 * The image this backend was built in has no node / tsc; this file is delivered as source, written against the
 * reference's own types (`src/api/types.ts`), and has not been compiled here.
 */
import type { FHEEngine } from '../src/api/fhe-engine';
import type { Ciphertext, EvaluationKey, BootstrapKey, ProgressCallback } from '../src/api/types';
import { FHEError, FHEErrorCode } from '../src/api/types';
import type * as fheb from './fheb_addon';

type Addon = typeof fheb;

/** words of every live ciphertext, by handle */
class CiphertextStore {
  private next = 1n;
  private readonly words = new Map<bigint, BigUint64Array>();
  put(w: BigUint64Array): bigint {
    const h = this.next++;
    this.words.set(h, w);
    return h;
  }
  get(ct: Ciphertext): BigUint64Array {
    const w = this.words.get(ct.handle);
    if (!w) throw new FHEError('Unknown ciphertext handle', FHEErrorCode.INVALID_PARAMETERS);
    return w;
  }
  drop(ct: Ciphertext): void {
    this.words.delete(ct.handle);
  }
}

export interface B200Keys {
  /** FHEE container of the evaluation key (KeySerializer::serialize_eval_key bytes) */
  evalKeyWire?: Buffer;
  /** bootstrapping key words [n][(k+1)L][k+1][N] and its shape */
  bootstrap?: { bsk: BigUint64Array; lweDimension: number; glweDimension: number; baseLog: number; level: number };
  /** test polynomial of programmableBootstrap (N words); default = the engine's default LUT */
  testPoly?: BigUint64Array;
}

export interface B200Engine extends FHEEngine {
  importCiphertext(words: BigUint64Array, meta: Omit<Ciphertext, '__brand' | 'handle'>): Ciphertext;
  exportCiphertext(ct: Ciphertext): BigUint64Array;
  /** EncryptionEngine::tally_votes over raw ballots [count][2][N] (e.g. the output of ingestBallots) */
  tallyBallots(ballots: BigUint64Array, keyId: bigint, noiseBudget: number): Ciphertext;
}

export function wireB200(base: FHEEngine, addon: Addon, keys: B200Keys = {}): B200Engine {
  addon.initialize();
  addon.setDevices(); // one Node process, every visible GPU
  const params = base.getParams();
  const N = params.polyDegree;
  const q = BigInt(params.moduli[0]!);
  const store = new CiphertextStore();
  let boot: fheb.BootstrapEngine | undefined;

  const make = (w: BigUint64Array, m: Omit<Ciphertext, '__brand' | 'handle'>): Ciphertext => ({ __brand: 'Ciphertext', handle: store.put(w), ...m });
  const sameKey = (a: Ciphertext, b: Ciphertext): void => {
    if (a.keyId !== b.keyId) throw new FHEError('Key mismatch', FHEErrorCode.KEY_MISMATCH); // fhe-engine.ts:211-213
  };

  const engine = Object.create(base) as B200Engine;

  engine.importCiphertext = (words, meta) => make(words, meta);
  engine.exportCiphertext = (ct) => store.get(ct);

  // EncryptionEngine::add (encryption.cpp:594-618): component-wise sum, noise budget min - 1
  engine.add = async (a, b) => {
    sameKey(a, b);
    return make(addon.modAddBatch(store.get(a), store.get(b), q),
                { keyId: a.keyId, noiseBudget: Math.min(a.noiseBudget, b.noiseBudget) - 1, isNtt: a.isNtt, degree: Math.max(a.degree, b.degree) });
  };
  engine.subtract = async (a, b) => {
    sameKey(a, b);
    return make(addon.modSubBatch(store.get(a), store.get(b), q),
                { keyId: a.keyId, noiseBudget: Math.min(a.noiseBudget, b.noiseBudget) - 1, isNtt: a.isNtt, degree: Math.max(a.degree, b.degree) });
  };

  // batchAdd (fhe-engine.ts:251-267 folds with add; the words are those of EncryptionEngine::batch_add for any grouping):
  // ONE tally call over the concatenated ciphertexts instead of count-1 additions
  engine.batchAdd = async (cts: Ciphertext[], progress?: ProgressCallback) => {
    if (cts.length === 0) throw new FHEError('Empty array', FHEErrorCode.INVALID_PARAMETERS);
    const start = Date.now();
    let budget = cts[0]!.noiseBudget;
    const all = new BigUint64Array(cts.length * 2 * N);
    cts.forEach((ct, i) => {
      if (i > 0) {
        sameKey(cts[0]!, ct);
        budget = Math.min(budget, ct.noiseBudget) - 1; // the reference's fold: min - 1 per addition
      }
      all.set(store.get(ct), i * 2 * N);
    });
    const words = cts.length === 1 ? store.get(cts[0]!) : addon.tallyVotes(all, N, q);
    progress?.({ stage: 'batch_add', current: cts.length, total: cts.length, elapsedMs: Date.now() - start, progressPercent: 100 });
    return make(words, { keyId: cts[0]!.keyId, noiseBudget: budget, isNtt: cts[0]!.isNtt, degree: cts[0]!.degree });
  };
  engine.tallyBallots = (ballots, keyId, noiseBudget) => {
    const count = ballots.length / (2 * N);
    // EncryptionEngine::tally_votes = batch_add_tree: min - 1 per level (encryption.cpp:1413,1437)
    return make(addon.tallyVotes(ballots, N, q), { keyId, noiseBudget: noiseBudget - Math.ceil(Math.log2(Math.max(count, 1))), isNtt: false, degree: 1 });
  };

  // multiply + relinearize (encryption.cpp:737-798, 904-993).  The addon fuses both around a pre-transformed key.
  engine.multiplyRelin = async (a, b, _ek: EvaluationKey) => {
    sameKey(a, b);
    if (!keys.evalKeyWire) throw new FHEError('No evaluation key uploaded', FHEErrorCode.INVALID_PARAMETERS);
    return make(addon.multiplyRelinearize(N, q, keys.evalKeyWire, store.get(a), store.get(b)),
                { keyId: a.keyId, noiseBudget: Math.min(a.noiseBudget, b.noiseBudget) / 2, isNtt: a.isNtt, degree: 1 });
  };
  engine.squareRelin = async (a, ek) => engine.multiplyRelin(a, a, ek);

  // bootstrap / programmableBootstrap (bootstrap_engine.cpp:684-723): LWE words [n+1] in, [k*N+1] (or [n_out+1]) out
  const bootstrapWith = async (ct: Ciphertext, _bk: BootstrapKey, testPoly?: BigUint64Array): Promise<Ciphertext> => {
    const k = keys.bootstrap;
    if (!k) throw new FHEError('No bootstrapping key uploaded', FHEErrorCode.INVALID_PARAMETERS);
    boot ??= new addon.BootstrapEngine(N, q, k.lweDimension, k.glweDimension, k.baseLog, k.level, k.bsk);
    const tp = testPoly ?? keys.testPoly;
    if (!tp) throw new FHEError('No test polynomial', FHEErrorCode.INVALID_PARAMETERS);
    return make(boot.bootstrapBatch(store.get(ct), tp), { keyId: ct.keyId, noiseBudget: params.noiseBudget, isNtt: ct.isNtt, degree: ct.degree });
  };
  engine.bootstrap = (ct, bk) => bootstrapWith(ct, bk);
  engine.programmableBootstrap = (ct, bk, lut) => bootstrapWith(ct, bk, lut.length === N ? BigUint64Array.from(lut) : undefined);

  engine.getHardwareCapabilities = () => ({ ...base.getHardwareCapabilities(), ...addon.detectHardware() });
  return engine;
}

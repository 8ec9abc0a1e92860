/* Types of the Node-API addon in this directory: the reference's index.d.ts (lines 14-44, unchanged) plus the bulk
 * entry points FHEEngineImpl needs (src/api/fhe-engine.ts:209-321).  Polynomial data are BigUint64Array, row-major
 * [batch][N]; 64-bit scalars are bigint. */
export declare function initialize(): void
export interface HardwareCapabilities {
  hasSme: boolean
  hasMetal: boolean
  hasNeon: boolean
  hasAmx: boolean
  metalGpuCores: number
  unifiedMemorySize: number
}
export declare function detectHardware(): HardwareCapabilities
export declare function version(): string
export declare class ModularArithmetic {
  constructor(modulus: number)
  montgomeryMul(a: number, b: number): number
  modAdd(a: number, b: number): number
  modSub(a: number, b: number): number
  toMontgomery(a: number): number
  fromMontgomery(a: number): number
  getModulus(): number
}
/* ---- bulk additions ---- */
export declare class NttProcessor {
  constructor(degree: number, modulus: bigint)
  /** in place; returns its argument */
  forwardBatch(coeffs: BigUint64Array): BigUint64Array
  inverseBatch(coeffs: BigUint64Array): BigUint64Array
  /** PolynomialRing::multiply over a batch, coefficient form in and out */
  polymulBatch(a: BigUint64Array, b: BigUint64Array): BigUint64Array
}
export declare class BootstrapEngine {
  /** bsk: [lweDimension][(k+1)*level][k+1][degree] coefficient-form words */
  constructor(degree: number, modulus: bigint, lweDimension: number, glweDimension: number, baseLog: number, level: number, bsk: BigUint64Array)
  /** lwe: [batch][n+1]; returns [batch][k*degree+1] */
  bootstrapBatch(lwe: BigUint64Array, testPoly: BigUint64Array): BigUint64Array
}
/** ballots: [count][2][degree]; returns [2][degree] */
export declare function tallyVotes(ballots: BigUint64Array, degree: number, modulus: bigint): BigUint64Array
/** FHEV records back to back; status[i]: 0 ok, 1 too small, 2 bad magic, 3 checksum, 4 shape */
export declare function ingestBallots(wire: Buffer, count: number, numChoices: number, degree: number, modulus: bigint):
  { ciphertexts: BigUint64Array, status: Uint8Array, accepted: number }
/** EncryptionEngine::multiply + relinearize; evalKeyWire: an FHEE container */
export declare function multiplyRelinearize(degree: number, modulus: bigint, evalKeyWire: Buffer, ct1: BigUint64Array, ct2: BigUint64Array): BigUint64Array

/*
 * JNI veneer over librapid_b200.so (include/rapid_b200.h).  UNCOMPILED in this repository's build image (no JDK);
 * it is the binding a Rapid maintainer adds: one static native per C entry point, handles as long, arrays as direct
 * ByteBuffers / primitive arrays.  Every method returns the C status code (0 ok, <0 error); lastError() is the message.
 */
package com.vrg.rapid.gpu;

import java.nio.ByteBuffer;

public final class Native {
    static {
        System.loadLibrary("rapid_jni");     // java/jni/rapid_jni.c, linked against librapid_b200.so
    }

    private Native() {
    }

    public static native String lastError();

    // ---- MembershipView (com.vrg.rapid.MembershipView) ----
    /** rapid_view_create: hostBytes = hostnames concatenated, hostOff[n+1]; returns handle or 0 (see lastError). */
    public static native long viewCreate(int k, long n, byte[] hostBytes, int[] hostOff, int[] port, int device);
    public static native int viewDestroy(long view);
    public static native int viewRing(long view, int ring, int[] outIds);
    public static native int viewObservers(long view, int node, int[] outK);          // returns count or <0
    public static native int viewSubjects(long view, int node, int[] outK);
    public static native int viewExpectedObservers(long view, byte[] host, int port, int[] outK);
    public static native int viewRingNumbers(long view, int observer, int subject);   // bitmask or <0
    public static native int viewConfigId(long view, long[] idHigh, long[] idLow, long[] out1);
    public static native int viewRegisterJoiners(long view, byte[] hostBytes, int[] hostOff, int[] port);  // first id

    // ---- MultiNodeCutDetector / alert-batch handler ----
    public static native long cdCreate(long view, int h, int l, long receivers, long receiverBegin, int modeFlags,
                                       long maxSubjects);
    public static native int cdDestroy(long cd);
    /** rapid_cd_apply_batch with direct buffers: dst int32[n], ring uint8[n], status uint8[n]; outputs may be null. */
    public static native int cdApplyBatch(long cd, long cfgId, long nCells, ByteBuffer dst, ByteBuffer ring,
                                          ByteBuffer status, ByteBuffer cellCfg, int deliveryFlags, ByteBuffer blocked,
                                          ByteBuffer bitmap, long permSeed, ByteBuffer outHash, ByteBuffer outHash2,
                                          ByteBuffer outLen, ByteBuffer outAnnounced);
    public static native int cdGetProposal(long cd, long receiver, int[] outIds);      // returns length or <0
    public static native int cdAggregate(long cd, int[] dst, byte[] ring, byte[] status, long receiver, int[] outIds);
    public static native int cdInvalidate(long cd, long receiver, int[] outIds);
    public static native int cdNumProposals(long cd, long receiver);
    public static native int cdClear(long cd);

    // ---- FastPaxos fast round ----
    public static native long fpCreate(long cfgId, long membershipSize, long senderCapacity, int device);
    public static native int fpDestroy(long fp);
    public static native int fpReset(long fp, long cfgId, long membershipSize);
    /** out6 = {decided, hashLo.. } see rapid_fp_tally; returns status */
    public static native int fpTally(long fp, int[] sender, long[] voteCfg, long[] hash, long[] hash2, int[] len,
                                     long[] out6);
    public static native int fpTallyCd(long fp, long cd, long comm, long[] out6);

    public static native long[] proposalFingerprint(int[] ids);
}

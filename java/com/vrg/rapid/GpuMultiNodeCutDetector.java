/*
 * Drop-in for com.vrg.rapid.MultiNodeCutDetector (same package: the original is package-private and final, and is
 * constructed directly at Cluster.java:265-267 / :458-460).  UNCOMPILED here (no JDK in the build image).
 *
 * Seam 1 of INTEGRATION.md: one detector per process, RAW handle, R = 1.  Endpoints are mapped to the int32 ids of the
 * GpuMembershipView of the current configuration; joiners (UP alerts about non-members) are registered on first sight.
 */
package com.vrg.rapid;

import com.vrg.rapid.gpu.Native;
import com.vrg.rapid.pb.AlertMessage;
import com.vrg.rapid.pb.EdgeStatus;
import com.vrg.rapid.pb.Endpoint;

import java.util.ArrayList;
import java.util.Collections;
import java.util.List;

final class GpuMultiNodeCutDetector {
    private static final int RAPID_CD_RAW = 1;
    private final GpuMembershipView view;
    private final long handle;
    private final int[] scratch = new int[1 << 16];

    GpuMultiNodeCutDetector(final GpuMembershipView view, final int K, final int H, final int L) {
        if (H > K || L > H || K < 3 || L <= 0 || H <= 0) {                 // MultiNodeCutDetector.java:52-55
            throw new IllegalArgumentException("Arguments do not satisfy K > H >= L >= 0:"
                    + " (K: " + K + ", H: " + H + ", L: " + L);
        }
        this.view = view;
        this.handle = Native.cdCreate(view.handle(), H, L, 1, 0, RAPID_CD_RAW, 0);
        if (handle == 0) {
            throw new IllegalStateException(Native.lastError());
        }
    }

    int getNumProposals() {                                                // :62-66
        return Native.cdNumProposals(handle, 0);
    }

    List<Endpoint> aggregateForProposal(final AlertMessage msg) {          // :76-82
        final int n = msg.getRingNumberCount();
        final int dstId = view.idOf(msg.getEdgeDst(), msg.getEdgeStatus() == EdgeStatus.UP);
        final int[] dst = new int[n];
        final byte[] ring = new byte[n];
        final byte[] status = new byte[n];
        for (int i = 0; i < n; i++) {
            dst[i] = dstId;
            ring[i] = (byte) msg.getRingNumber(i);
            status[i] = (byte) msg.getEdgeStatusValue();
        }
        final int len = Native.cdAggregate(handle, dst, ring, status, 0, scratch);
        return toEndpoints(len);
    }

    List<Endpoint> invalidateFailingEdges(final MembershipView ignored) {  // :137-164 (the GPU view is already bound)
        return toEndpoints(Native.cdInvalidate(handle, 0, scratch));
    }

    void clear() {                                                         // :169-178
        Native.cdClear(handle);
    }

    private List<Endpoint> toEndpoints(final int len) {
        if (len < 0) {
            throw new IllegalStateException(Native.lastError());
        }
        if (len == 0) {
            return Collections.emptyList();
        }
        final List<Endpoint> out = new ArrayList<>(len);
        for (int i = 0; i < len; i++) {
            out.add(view.endpointOf(scratch[i]));
        }
        return out;
    }
}

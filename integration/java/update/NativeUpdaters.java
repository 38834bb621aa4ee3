package update;

import nativeps.PsNative;

/**
 * How the reference's updater objects reach the native store.  update/AdamUpdater.java:43,50 and
 * update/FtrlUpdater.java:37,44 stay as they are (their update() is never called for keys the native step owns);
 * what matters is their getName() string — "adam@alfa:..@beta1:..@" (AdamUpdater.java:84-88; FtrlUpdater.java:78-80 also
 * says "adam@", a copy-paste the reference keeps unique through the hyper-parameters) — which PServer uses as the registry
 * key (CTR.java:77-80).  ps_updater_parse turns that string into the {kind, p0..p3} spec that modelCreate / fcCreate take,
 * so `updaters.put("emF", ftrl)` (KVStore.java:244-248 prefix rule) becomes modelCreate(..., specOf(ftrl), ...).
 * SOURCE ONLY: no JDK in the build image.
 */
public final class NativeUpdaters {
	private NativeUpdaters() {}
	/** {kind, p0, p1, p2, p3} for an Updater, through its name as the reference registers it. */
	public static float[] specOf(Updater u) { return PsNative.updaterParse(u.getName()); }
}

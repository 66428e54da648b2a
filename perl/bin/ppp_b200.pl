#!/usr/bin/env perl
# perl/bin/ppp_b200.pl -- command-line drop-in for ProPhylo's bin/ppp.pl on the B200 path.
#
# Same options, same database layout (<db>/blast/<taxon>/<gi>.bla), same stdout table as the reference's server mode
# (/root/reference/bin/ppp.pl:300-515): header `Locus #Yes #Total Depth Prob Score Desc`, rows sorted by raw score,
# `%.3f` Prob and log10 Score, the optional slope marker line (-m).  What changes is the middle: instead of writing
# chunk files of 50 cache paths and forking `perl ppp.pl --client` workers that parse every cache and sweep every gene
# in Perl (:354-430, :519-593, :895-1017), the whole genome is ONE GPU call: the `.bla` caches are parsed natively
# (ppp_read_bla) and every gene is scored by libppp_b200.so through the XS binding (PhyloProf::B200).
# There is no CPU fallback.  Written from scratch against the reference's observable behaviour.
#
#   perl ppp_b200.pl -d <db dir> -t <taxon> -p <profile> -w <workspace> [--prob P] [-m slope] [--dup] [--device N | --devices 0-7]
#
# -l/--level <rank> (bin/ppp.pl:532-547, 566-568): the query profile's taxa and every taxon of every hit list are collapsed
# to that rank on the host before packing (Profile.pm:516-549, ppp.pm:328-370) through SeqToolBox::Taxonomy, which must be
# loadable (the reference's lib/ on PERL5LIB and its taxonomy database in place); the caches are then parsed in Perl.
# --dup (bin/ppp.pl:282 -> ppp.pm:189-200) is packed natively (PPP_BLA_DUP) and scored on the device.
# --threads/--serial/--keep/-c/-o/-r/-n are accepted and ignored (there are no workers, temporary files or grid jobs).
use strict;
use warnings;
use Getopt::Long;
use File::Spec;
use FindBin;
use lib "$FindBin::Bin/../lib";
use PhyloProf::B200;

my ( $help, $db_dir, $workspace, $taxon, $profile, $level, $probability, $slope, $duplicate, $device, $devices ) = ( 0, '', '', 0, '', '', 0, 0, 0, 0, '' );
my %ignored;
GetOptions(
    'help|h'      => \$help,      'db|d=s'   => \$db_dir, 'workspace|w=s' => \$workspace,   'taxon|t=i' => \$taxon,
    'profile|p=s' => \$profile,   'level|l=s' => \$level, 'prob=s'        => \$probability, 'slope|m=s' => \$slope,
    'dup'         => \$duplicate, 'device=i' => \$device, 'devices=s' => \$devices,
    'project|r=i' => \$ignored{r}, 'nodes|n=i' => \$ignored{n}, 'client' => \$ignored{client}, 'copy|c' => \$ignored{c},
    'output|o=s'  => \$ignored{o}, 'serial'    => \$ignored{serial}, 'threads=i' => \$ignored{threads}, 'gi|g=i' => \$ignored{g},
    'start-depth|s=i' => \$ignored{s}, 'end-depth|e=i' => \$ignored{e}, 'skip|j=i' => \$ignored{j}, 'keep|k' => \$ignored{k},
) or die "usage: ppp_b200.pl -d <db> -t <taxon> -p <profile> -w <workspace> [--prob P] [-m slope]\n";
die "usage: ppp_b200.pl -d <db> -t <taxon> -p <profile> -w <workspace> [--prob P] [-m slope]\n" if $help || !$taxon || !$profile;
die "File $profile not found\n" unless -s $profile;
die "Workspace missing\n"       unless $workspace;
die "ppp_b200.pl: --client is the reference's worker mode; this driver has no workers\n" if $ignored{client};
my $collapse = sub { $_[0] };
if ($level) {
    eval { require SeqToolBox::Taxonomy; 1 } or die "ppp_b200.pl: -l/--level needs SeqToolBox::Taxonomy on \@INC: $@";
    my $tax = SeqToolBox::Taxonomy->new();
    my %memo;
    $collapse = sub {    # ppp.pm:328-370 / Profile.pm:516-549: the rank, else genus, else the raw id
        my $t = shift;
        return $t unless $t;
        $memo{$t} = $tax->collapse_taxon( $t, $level ) || $tax->collapse_taxon( $t, "genus" ) || $t unless exists $memo{$t};
        return $memo{$t};
    };
}

# ---- the gene caches of the genome (bin/ppp.pl:955-1017 lists them the same way)
my $blast_dir = File::Spec->catdir( $db_dir, "blast", $taxon );
opendir( my $dh, $blast_dir ) or die "Can't find blast result. Did you setup the database properly?\n";
my ( @loci, @paths );
foreach my $file ( sort readdir($dh) ) {
    next unless $file =~ /(\S+)\.bla$/;
    push @loci,  $1;
    push @paths, File::Spec->catfile( $blast_dir, $file );
}
closedir($dh);
die "Can't find blast result. Did you setup the database properly?\n" unless @paths;

# ---- the query profile: two tab columns taxon, 0/1; '#' comments; on a repeated taxon a non-zero value is never
# overwritten (lib/PhyloProf/Profile.pm:433-514)
my %q;
open( my $pf, "<", $profile ) or die "Can't open $profile\n";
while ( my $line = <$pf> ) {
    chomp $line;
    next if $line =~ /^\#/;
    my @f = split( /\t/, $line );
    die "$profile must contain only two fields\n" if scalar(@f) != 2;
    my $taxon = $collapse->( $f[0] );
    next unless $taxon;                                  # a false taxon ('' or '0') is skipped
    next if exists $q{$taxon} && $q{$taxon} != 0;
    $q{$taxon} = $f[1];
}
close($pf);
my $yes   = grep { $_ != 0 } values %q;
my $total = scalar keys %q;
die "Illegal division by zero\n" unless $total || $probability;
my $p = $probability ? $probability + 0 : $yes / $total;    # ppp.pm:138-142, bin/ppp.pl:559-572

# ---- one GPU call for the whole genome
my @universe = sort keys %q;
my $gpu      = $devices ne '' ? PhyloProf::B200->new( -devices => $devices ) : PhyloProf::B200->new( -device => $device );
my $hits;
if ($level) {
    # rank collapse: the hit lists are read here (bin/ppp.pl:606-673 + BlastResult.pm:192-237: subjects in first-occurrence
    # order, a subject's last taxon wins, false taxa dropped) and collapsed taxon by taxon before the ranked-list entry
    my @lists;
    foreach my $path (@paths) {
        open( my $bf, "<", $path ) or die "Can't open $path\n";
        my ( @order, %taxon );
        while ( my $line = <$bf> ) {
            chomp $line;
            my @f = split( /\t/, $line );
            my $subject = defined $f[1] ? $f[1] : '';
            push @order, $subject unless exists $taxon{$subject};
            $taxon{$subject} = $f[3];
        }
        close($bf);
        push @lists, [ map { $collapse->($_) } grep { $_ } map { $taxon{$_} } @order ];
    }
    $hits = $gpu->score_ranked( \@universe, [ \%q ], \@lists, -probs => [$p], -filter => 'all', -dup => $duplicate );
} else {
    $hits = $gpu->score_bla_files( \@universe, [ \%q ], \@paths, -probs => [$p], -filter => 'all', -dup => $duplicate );
}

# ---- the table (bin/ppp.pl:443-506).  The reference's clients print score and p with Perl's default %.15g
# stringification and the server reads them back: reproduce that rounding.
my ( %score, %yes, %number, %depth );
foreach my $h (@$hits) {
    my $locus = $loci[ $h->[1] ];
    $score{$locus}  = 0 + sprintf( "%.15g", $h->[2] );
    $yes{$locus}    = $h->[3];
    $number{$locus} = $h->[4];
    $depth{$locus}  = $h->[5];
}
my $prob = 0 + sprintf( "%.15g", $p );
my $sth;
my $desc_db = File::Spec->catfile( $db_dir, "desc", "gi_desc.sqlite3" );
if ( -s $desc_db && eval { require DBI; 1 } ) {    # descriptions are optional: the column stays empty without the database
    my $dbh = eval { DBI->connect( "dbi:SQLite:dbname=$desc_db", "", "", { RaiseError => 1, AutoCommit => 0 } ) };
    $sth = eval { $dbh->prepare('select desc from gi_desc where gi =?') } if $dbh;
}
print "Locus\t\#Yes\t#Total\tDepth\tProb\tScore\tDesc\n";
my ( $last_score, $cutoff_found );
foreach my $s ( sort { $score{$a} <=> $score{$b} || $a cmp $b } keys %score ) {
    die "Can't take log of 0\n" if $score{$s} == 0;    # as the reference does (log(0) is fatal, bin/ppp.pl:476)
    my $log_score = sprintf( "%.3f", log( $score{$s} ) / log(10) );
    $last_score = $log_score unless $last_score;
    if ( $slope && $last_score && !$cutoff_found ) {
        my $pl = $last_score < 0 ? -$last_score : $last_score;
        my $ps = $log_score < 0  ? -$log_score  : $log_score;
        if ( ( ( $pl - $ps ) / $pl ) * 100 > $slope ) {
            print "-------------------------------------------\n";
            $cutoff_found = 1;
        }
    }
    my $desc = "";
    if ($sth) { $sth->execute($s); my @row = $sth->fetchrow_array(); $desc = $row[0] if $row[0]; }
    $log_score = "0.000" if $log_score eq "-0.000";    # handle_negative_zero, bin/ppp.pl:595-604
    print "$s\t$yes{$s}\t$number{$s}\t$depth{$s}\t", sprintf( "%.3f", $prob ), "\t$log_score\t$desc\n";
}
exit(0);
